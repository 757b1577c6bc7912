#include <slam/moving_laser_scan.hpp>
#include <lcmtypes/lidar_t.hpp>
#include <lcmtypes/pose_xyt_t.hpp>
#include <cmath>

namespace {

// Arithmetic contract of the reference: float pose fields, double interpolation, results
// rounded to float; the same sequence the sensor kernel evaluates per particle (csrc/mcl_device.cuh).
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

MovingLaserScan::MovingLaserScan(const lidar_t& scan, const pose_xyt_t& beginPose, const pose_xyt_t& endPose,
                                 int rayStride)
{
    if (scan.num_ranges <= 0) return;
    if (rayStride < 1) rayStride = 1;
    const bool moving = beginPose.utime != endPose.utime;
    const double span = static_cast<double>(endPose.utime - beginPose.utime);
    const double dx = static_cast<double>(endPose.x - beginPose.x);      // float difference, widened
    const double dy = static_cast<double>(endPose.y - beginPose.y);
    const double dth = foldToPi(static_cast<double>(endPose.theta) - static_cast<double>(beginPose.theta));
    rays_.reserve(static_cast<std::size_t>(scan.num_ranges / rayStride + 1));
    for (int n = 0; n < scan.num_ranges; n += rayStride) {
        if (!(scan.ranges[n] > 0.15f)) continue;
        adjusted_ray_t ray;
        float heading;
        if (moving) {
            const double ratio = static_cast<double>(scan.times[n] - beginPose.utime) / span;
            ray.origin.x = static_cast<float>(static_cast<double>(beginPose.x) + dx * ratio);
            ray.origin.y = static_cast<float>(static_cast<double>(beginPose.y) + dy * ratio);
            heading = static_cast<float>(foldToPi(static_cast<double>(beginPose.theta) + dth * ratio));
        } else {
            ray.origin.x = endPose.x;
            ray.origin.y = endPose.y;
            heading = endPose.theta;
        }
        ray.range = scan.ranges[n];
        ray.theta = wrapToPi(heading - scan.thetas[n]);     // the lidar reports clockwise angles
        rays_.push_back(ray);
    }
}
