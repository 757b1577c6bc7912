// Oracle-only stand-in for the lcm-gen output of lcmtypes/pose_xyt_t.lcm:1-8 (lcm-gen is absent in this image).
// Members are zero-initialised so the reference's estimatePosteriorPose (particle_filter.cpp:146) starts from 0.
#ifndef ORACLE_SHIM_POSE_XYT_T_HPP
#define ORACLE_SHIM_POSE_XYT_T_HPP
#include <cstdint>
class pose_xyt_t
{
public:
    int64_t utime = 0;
    float x = 0.0f;
    float y = 0.0f;
    float theta = 0.0f;
};
#endif
