// Oracle-only stand-in for lcmtypes/particle_t.lcm:4-9.
#ifndef ORACLE_SHIM_PARTICLE_T_HPP
#define ORACLE_SHIM_PARTICLE_T_HPP
#include "pose_xyt_t.hpp"
class particle_t
{
public:
    pose_xyt_t pose;
    pose_xyt_t parent_pose;
    double weight;
};
#endif
