// DeviceFilter -- RAII owner of one mcl_engine plus the OccupancyGrid device-mirror bookkeeping.  Used by the host
// classes (ParticleFilter, SensorModel, ActionModel) to reach the CUDA engine through the C ABI in mcl_cuda.h.
#ifndef B200_SLAM_CUDA_DEVICE_FILTER_HPP
#define B200_SLAM_CUDA_DEVICE_FILTER_HPP

#include <slam/cuda/mcl_cuda.h>
#include <slam/occupancy_grid.hpp>
#include <stdexcept>
#include <string>

class lidar_t;

namespace b200 {

/// Thrown when the engine reports an error; what() carries mcl_last_error().  There is no CPU fallback to fall to.
class EngineError : public std::runtime_error
{
public:
    EngineError(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
    int code(void) const { return code_; }
private:
    int code_;
};

class DeviceFilter
{
public:
    DeviceFilter(int64_t numParticles, int device = 0, const mcl_params* params = nullptr);
    ~DeviceFilter(void);
    DeviceFilter(const DeviceFilter&) = delete;
    DeviceFilter& operator=(const DeviceFilter&) = delete;

    mcl_engine* engine(void) const { return engine_; }
    int64_t numParticles(void) const { return numParticles_; }

    /// Brings the device mirror up to date with `map`: full upload when the grid's (process-wide unique) generation
    /// differs from the one mirrored, otherwise only the rectangle of cells written since the write sequence this
    /// mirror last saw.  Any number of mirrors can follow one grid.
    void syncMap(const OccupancyGrid& map);
    /// After cells were read back FROM this mirror into `map` (OccupancyGrid::noteExternalWrite returned seq): this
    /// mirror already holds them, provided it was in sync right before.
    void noteMirrorIsAheadOf(const OccupancyGrid& map, uint64_t seq);

    void check(int rc) const;   ///< throws EngineError on rc != 0

private:
    mcl_engine* engine_;
    int64_t numParticles_;
    uint64_t mirroredGeneration_;
    uint64_t mirroredSeq_;
};

/// Environment knobs shared by the host classes: B200_MCL_DEVICE (CUDA ordinal, default 0); B200_MCL_LEGACY_UTIME=1
/// selects mcl_params::legacy_equal_utime, i.e. the UNMODIFIED reference's behaviour (ActionModel::utime_ is never
/// assigned there, so every ray starts at the current pose); the default is the evident intent (per-ray interpolation).
int defaultDevice(void);
bool defaultLegacyEqualUtime(void);

}  // namespace b200

#endif
