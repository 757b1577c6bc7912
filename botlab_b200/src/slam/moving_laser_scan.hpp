// MovingLaserScan -- a scan whose rays each start from the pose the robot had when that ray was measured, by linear
// interpolation between the scan's begin and end poses.  Same public interface as the reference's
// src/slam/moving_laser_scan.hpp:14-57.  The particle filter no longer builds one of these per particle -- the
// sensor kernel interpolates rays in registers -- but Mapping and tests still use the host object.
#ifndef B200_SLAM_MOVING_LASER_SCAN_HPP
#define B200_SLAM_MOVING_LASER_SCAN_HPP

#include <common/point.hpp>
#include <cstddef>
#include <vector>

class lidar_t;
class pose_xyt_t;

struct adjusted_ray_t
{
    Point<float> origin;   ///< robot position when the ray was measured
    float range;           ///< measured range (m)
    float theta;           ///< ray heading in the global frame
};

class MovingLaserScan
{
public:
    typedef std::vector<adjusted_ray_t>::const_iterator Iter;

    /// Rays with range <= 0.15 m (inside the robot) are dropped; rayStride < 1 is treated as 1.
    MovingLaserScan(const lidar_t& scan, const pose_xyt_t& beginPose, const pose_xyt_t& endPose, int rayStride = 1);

    std::size_t size(void) const { return rays_.size(); }
    Iter begin(void) const { return rays_.begin(); }
    Iter end(void) const { return rays_.end(); }
    adjusted_ray_t at(int index) const { return rays_.at(index); }
    const adjusted_ray_t& operator[](int index) const { return rays_[index]; }

private:
    std::vector<adjusted_ray_t> rays_;
};

#endif
