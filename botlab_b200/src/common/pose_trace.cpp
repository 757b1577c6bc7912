#include <common/pose_trace.hpp>
#include <cmath>
#include <iostream>

namespace {
double foldToPi(double a)
{
    if (std::fabs(a) > M_PI) a -= (a > 0) ? 2 * M_PI : -2 * M_PI;
    return a;
}
float wrapToPi(float a)
{
    while (static_cast<double>(a) < -M_PI) a = static_cast<float>(static_cast<double>(a) + 2.0 * M_PI);
    while (static_cast<double>(a) > M_PI) a = static_cast<float>(static_cast<double>(a) - 2.0 * M_PI);
    return a;
}
}  // namespace

pose_xyt_t interpolate_pose_by_time(int64_t time, const pose_xyt_t& before, const pose_xyt_t& after)
{
    pose_xyt_t out;
    if (before.utime == after.utime) {
        out = after;
        out.utime = time;
        return out;
    }
    const double ratio = static_cast<double>(time - before.utime) / static_cast<double>(after.utime - before.utime);
    const double xs = static_cast<double>(after.x - before.x) * ratio;
    const double ys = static_cast<double>(after.y - before.y) * ratio;
    const double ts = foldToPi(static_cast<double>(after.theta) - static_cast<double>(before.theta)) * ratio;
    out.utime = time;
    out.x = static_cast<float>(static_cast<double>(before.x) + xs);
    out.y = static_cast<float>(static_cast<double>(before.y) + ys);
    out.theta = static_cast<float>(foldToPi(static_cast<double>(before.theta) + ts));
    return out;
}

void PoseTrace::addPose(const pose_xyt_t& pose)
{
    // the reference passes every pose through its (identity, until setReferencePose) frame transform, which wraps theta
    pose_xyt_t p = pose;
    p.theta = wrapToPi(pose.theta);
    trace_.push_back(p);
}

pose_xyt_t PoseTrace::poseAt(int64_t time) const
{
    if (trace_.empty()) {
        std::cerr << "ERROR: PoseTrace::poseAt: no odometry measurements to interpolate.\n";
        return pose_xyt_t();
    }
    if (time < trace_.front().utime) {
        std::cerr << "ERROR: PoseTrace::poseAt: no odometry before " << time << ", returning the first pose.\n";
        return trace_.front();
    }
    if (time > trace_.back().utime) {
        std::cerr << "ERROR: PoseTrace::poseAt: no odometry after " << time << ", returning the last pose.\n";
        return trace_.back();
    }
    for (std::size_t i = 1; i < trace_.size(); ++i)
        if (trace_[i - 1].utime <= time && time <= trace_[i].utime) return interpolate_pose_by_time(time, trace_[i - 1], trace_[i]);
    return trace_.back();
}

bool PoseTrace::containsPoseAtTime(int64_t time) const
{
    return !trace_.empty() && trace_.front().utime <= time && time <= trace_.back().utime;
}
