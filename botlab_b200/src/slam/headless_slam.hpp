// HeadlessSLAM -- OccupancyGridSLAM's update loop (reference src/slam/slam.cpp:88-291) without LCM: the same queueing,
// the same readiness test, odometry sampled from a PoseTrace at the scan's last timestamp, poses initialised at the
// first scan, the 100-range sanity gate, localisation then mapping.  Messages are pushed by the caller in arrival
// order instead of arriving on LCM channels; poses come back through currentPose() instead of SLAM_POSE.
// This is the "headless replay" next row of SURVEY.md section 8f: it lets `slam --localization-only <map>` (BASELINE
// configs[0]) run end to end on the engine in an image that has no LCM.
#ifndef B200_SLAM_HEADLESS_SLAM_HPP
#define B200_SLAM_HEADLESS_SLAM_HPP

#include <common/pose_trace.hpp>
#include <lcmtypes/lidar_t.hpp>
#include <lcmtypes/pose_xyt_t.hpp>
#include <slam/mapping.hpp>
#include <slam/occupancy_grid.hpp>
#include <slam/particle_filter.hpp>
#include <deque>
#include <string>

class HeadlessSLAM
{
public:
    enum Mode { full_slam = 0, localization_only = 1, action_only = 2 };

    /// hit/miss odds default to slam_main.cpp:22-23 (4 and 1); maxLaserDistance to slam.cpp:24 (5 m).
    HeadlessSLAM(int numParticles, Mode mode, int8_t hitOdds = 4, int8_t missOdds = 1, float maxLaserDistance = 5.0f);

    /// --localization-only <file>: slam.cpp:36-47.
    bool loadMap(const std::string& filename);
    /// Same, from an in-memory grid message.
    void setMap(const occupancy_grid_t& grid);
    void setInitialPose(const pose_xyt_t& pose) { initialPose_ = pose; }

    void handleOdometry(const pose_xyt_t& odometry);        // slam.cpp:131-141
    void handleLaser(const lidar_t& scan);                  // slam.cpp:90-128
    bool isReadyToUpdate(void) const;                       // slam.cpp:163-188
    /// One runSLAMIteration (slam.cpp:191-207).  Returns false if the scan failed the sanity gate.
    bool runSLAMIteration(void);
    /// Drains every scan that is ready; returns how many iterations ran.
    int spin(void);

    const pose_xyt_t& currentPose(void) const { return currentPose_; }
    const OccupancyGrid& map(void) const { return map_; }
    ParticleFilter& filter(void) { return filter_; }
    bool haveInitializedPoses(void) const { return haveInitializedPoses_; }
    int iterations(void) const { return iterations_; }
    int ignoredScans(void) const { return numIgnoredScans_; }
    /// Run Mapping::updateMap on the filter's device mirror of the map instead of the host loops (identical cells).
    void setDeviceMapping(bool on) { mapper_.useDeviceMirror(on ? &filter_ : nullptr); }

    /// Called right after initializeFilterAtPose on the first iteration (tests plant a deterministic cloud here).
    void (*onFilterInitialized)(HeadlessSLAM&, void*) = nullptr;
    /// Called before each localisation update with the iteration index (tests inject recorded action draws here).
    void (*beforeLocalization)(HeadlessSLAM&, int, void*) = nullptr;
    void* hookArg = nullptr;

private:
    Mode mode_;
    bool haveInitializedPoses_;
    bool haveMap_;
    int numIgnoredScans_;
    int iterations_;
    std::deque<lidar_t> incomingScans_;
    PoseTrace odometryPoses_;
    lidar_t currentScan_;
    pose_xyt_t currentOdometry_;
    pose_xyt_t initialPose_, previousPose_, currentPose_;
    ParticleFilter filter_;
    OccupancyGrid map_;
    Mapping mapper_;
};

#endif
