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

    /// Brings the device mirror up to date with `map`: full upload when the grid's generation or geometry changed,
    /// otherwise only the rectangle of cells written since the last call.
    void syncMap(const OccupancyGrid& map);

    void check(int rc) const;   ///< throws EngineError on rc != 0

private:
    mcl_engine* engine_;
    int64_t numParticles_;
    const OccupancyGrid* mirrored_;
    uint64_t mirroredGeneration_;
};

/// Environment knobs shared by the host classes: B200_MCL_DEVICE (CUDA ordinal, default 0).
int defaultDevice(void);

}  // namespace b200

#endif
