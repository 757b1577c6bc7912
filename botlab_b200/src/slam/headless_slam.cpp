#include <slam/headless_slam.hpp>
#include <iostream>

HeadlessSLAM::HeadlessSLAM(int numParticles, Mode mode, int8_t hitOdds, int8_t missOdds, float maxLaserDistance)
: mode_(mode), haveInitializedPoses_(false), haveMap_(false), numIgnoredScans_(0), iterations_(0),
  filter_(numParticles), map_(10.0f, 10.0f, 0.05f), mapper_(maxLaserDistance, hitOdds, missOdds)
{
    currentOdometry_.utime = 0;
    currentScan_.utime = 0;
}

bool HeadlessSLAM::loadMap(const std::string& filename)
{
    haveMap_ = map_.loadFromFile(filename);
    return haveMap_;
}

void HeadlessSLAM::setMap(const occupancy_grid_t& grid)
{
    map_.fromLCM(grid);
    haveMap_ = true;
}

void HeadlessSLAM::handleOdometry(const pose_xyt_t& odometry)
{
    odometryPoses_.addPose(odometry);
}

// A scan is queued only once odometry older than its first ray exists.
void HeadlessSLAM::handleLaser(const lidar_t& scan)
{
    const bool haveOdom = !odometryPoses_.empty() && !scan.times.empty() &&
                          odometryPoses_.front().utime <= scan.times.front();
    if (haveOdom) {
        incomingScans_.push_back(scan);
        numIgnoredScans_ = 0;
    } else {
        ++numIgnoredScans_;
    }
}

bool HeadlessSLAM::isReadyToUpdate(void) const
{
    if (incomingScans_.empty()) return false;
    return odometryPoses_.containsPoseAtTime(incomingScans_.front().times.front());
}

bool HeadlessSLAM::runSLAMIteration(void)
{
    // copyDataForSLAMUpdate
    currentScan_ = incomingScans_.front();
    incomingScans_.pop_front();
    currentOdometry_ = odometryPoses_.poseAt(currentScan_.times.back());
    // initializePosesIfNeeded: poses carry the first scan's timestamps so the first MovingLaserScan interpolates
    if (!haveInitializedPoses_) {
        previousPose_ = initialPose_;
        previousPose_.utime = currentScan_.times.front();
        currentPose_ = previousPose_;
        currentPose_.utime = currentScan_.times.back();
        haveInitializedPoses_ = true;
        filter_.initializeFilterAtPose(previousPose_);
        if (onFilterInitialized) onFilterInitialized(*this, hookArg);
    }
    if (!(currentScan_.num_ranges > 100)) {                  // rplidar lost sync
        std::cerr << "ERROR: HeadlessSLAM: invalid laser scan with " << currentScan_.num_ranges << " ranges.\n";
        return false;
    }
    // updateLocalization
    if (haveMap_) {
        previousPose_ = currentPose_;
        if (beforeLocalization) beforeLocalization(*this, iterations_, hookArg);
        currentPose_ = (mode_ == action_only) ? filter_.updateFilterActionOnly(currentOdometry_)
                                              : filter_.updateFilter(currentOdometry_, currentScan_, map_);
    }
    // updateMap: the reference's mode test (slam.cpp:276) is always true, so the map is updated in every mode
    mapper_.updateMap(currentScan_, currentPose_, map_);
    haveMap_ = true;
    ++iterations_;
    return true;
}

int HeadlessSLAM::spin(void)
{
    int ran = 0;
    while (isReadyToUpdate()) {
        runSLAMIteration();
        ++ran;
    }
    return ran;
}
