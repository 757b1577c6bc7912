// PoseTrace -- time-ordered odometry poses with linear interpolation; the subset of the reference's
// src/common/pose_trace.hpp that OccupancyGridSLAM's update loop uses (slam.cpp:138-141,163-188,227).
#ifndef B200_COMMON_POSE_TRACE_HPP
#define B200_COMMON_POSE_TRACE_HPP

#include <lcmtypes/pose_xyt_t.hpp>
#include <cstdint>
#include <vector>

/// Linear interpolation between two poses by time with the reference's arithmetic (common/interpolation.hpp:24-50):
/// equal utimes return `after`; otherwise the ratio and steps are doubles, the results rounded to float.
pose_xyt_t interpolate_pose_by_time(int64_t time, const pose_xyt_t& before, const pose_xyt_t& after);

class PoseTrace
{
public:
    void addPose(const pose_xyt_t& pose);
    /// Pose at `time`; clamps (with a message on stderr) to the first/last pose outside the trace, like the reference.
    pose_xyt_t poseAt(int64_t time) const;
    bool containsPoseAtTime(int64_t time) const;
    bool empty(void) const { return trace_.empty(); }
    std::size_t size(void) const { return trace_.size(); }
    const pose_xyt_t& front(void) const { return trace_.front(); }
    const pose_xyt_t& back(void) const { return trace_.back(); }

private:
    std::vector<pose_xyt_t> trace_;
};

#endif
