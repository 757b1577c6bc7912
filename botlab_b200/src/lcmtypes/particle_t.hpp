// Stand-in for lcm-gen's particle_t (lcmtypes/particle_t.lcm:4-9): 56 bytes.
#ifndef B200_LCMTYPES_PARTICLE_T_HPP
#define B200_LCMTYPES_PARTICLE_T_HPP
#include <lcmtypes/pose_xyt_t.hpp>
class particle_t
{
public:
    pose_xyt_t pose;
    pose_xyt_t parent_pose;
    double weight = 0.0;
};
#endif
