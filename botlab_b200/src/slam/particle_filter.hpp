// ParticleFilter -- SIR particle filter with the reference's public interface (src/slam/particle_filter.hpp:38-77),
// so OccupancyGridSLAM (slam.cpp:246,259-265) compiles against it unchanged.  All particle state lives on the GPU in
// the engine behind slam/cuda/mcl_cuda.h; one updateFilter call = one mcl_update (resample -> action -> sensor ->
// normalise -> estimate on one CUDA stream); the host sends the scan and reads back one pose.
#ifndef B200_SLAM_PARTICLE_FILTER_HPP
#define B200_SLAM_PARTICLE_FILTER_HPP

#include <slam/action_model.hpp>
#include <slam/sensor_model.hpp>
#include <lcmtypes/particle_t.hpp>
#include <lcmtypes/particles_t.hpp>
#include <lcmtypes/pose_xyt_t.hpp>
#include <memory>

class lidar_t;
class OccupancyGrid;
namespace b200 { class DeviceFilter; }

class ParticleFilter
{
public:
    /// \pre numParticles > 1
    ParticleFilter(int numParticles);
    /// Extension: explicit engine parameters (e.g. legacy_equal_utime = 1 for the unmodified reference's de-facto
    /// behaviour); the one-argument form takes the defaults and the B200_MCL_LEGACY_UTIME environment knob.
    ParticleFilter(int numParticles, const mcl_params& params);
    ~ParticleFilter(void);

    /// Cloud ~ pose + N(0, 0.01) per coordinate, last particle exactly at pose, weights 1/N.
    void initializeFilterAtPose(const pose_xyt_t& pose);

    /// One filter update.  If the odometry shows no motion nothing changes and the previous estimate is returned
    /// with utime = odometry.utime.
    pose_xyt_t updateFilter(const pose_xyt_t& odometry, const lidar_t& laser, const OccupancyGrid& map);

    /// Action model only (no resampling, no reweighting); returns the odometry pose.
    pose_xyt_t updateFilterActionOnly(const pose_xyt_t& odometry);

    pose_xyt_t poseEstimate(void) const;

    /// The posterior cloud.  At most maxExportedParticles() are copied back (every k-th one), because SLAM_PARTICLES
    /// consumers draw them from stack arrays (botgui drawing_functions.cpp:123-125).
    particles_t particles(void) const;

    // ---- extensions (not in the reference) ----
    void setMaxExportedParticles(int64_t n) { maxExported_ = n; }
    int64_t maxExportedParticles(void) const { return maxExported_; }
    /// particles() of a cloud larger than maxExportedParticles(): false (default) = every k-th particle with its weight;
    /// true = a systematic weighted draw of maxExportedParticles() particles, each with weight 1/count.
    void setExportWeighted(bool on) { exportWeighted_ = on; }
    /// Seeds both the device Philox stream (used from the next initializeFilterAtPose) and libc rand().
    void setSeed(uint64_t seed);
    mcl_stats stats(void) const;
    /// Replaces the cloud (all particles must share pose.utime and parent_pose.utime).
    void setParticles(const particles_t& cloud);
    /// Test hook: N x (rot1, trans, rot2) action draws used by the NEXT update instead of the Philox stream
    /// (the reference's recorded std::mt19937 draws).  The pointer must stay valid until that update returns.
    void injectActionNoise(const float* draws3n) { injectedNoise_ = draws3n; }
    /// The engine behind this filter, for components that work on its map mirror (Mapping::useDeviceMirror).
    b200::DeviceFilter& device(void) { return *device_; }

private:
    int kNumParticles_;
    pose_xyt_t posteriorPose_;
    ActionModel actionModel_;
    std::unique_ptr<b200::DeviceFilter> device_;
    uint64_t seed_;
    int64_t maxExported_;
    bool exportWeighted_;
    const float* injectedNoise_;
};

#endif
