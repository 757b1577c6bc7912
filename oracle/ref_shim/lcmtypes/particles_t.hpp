// Oracle-only stand-in for lcmtypes/particles_t.lcm:2-8.
#ifndef ORACLE_SHIM_PARTICLES_T_HPP
#define ORACLE_SHIM_PARTICLES_T_HPP
#include <vector>
#include "particle_t.hpp"
class particles_t
{
public:
    int64_t utime;
    int32_t num_particles;
    std::vector<particle_t> particles;
};
#endif
