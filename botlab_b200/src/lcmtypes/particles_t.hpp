// Stand-in for lcm-gen's particles_t (lcmtypes/particles_t.lcm:2-8).
#ifndef B200_LCMTYPES_PARTICLES_T_HPP
#define B200_LCMTYPES_PARTICLES_T_HPP
#include <vector>
#include <lcmtypes/particle_t.hpp>
class particles_t
{
public:
    int64_t utime = 0;
    int32_t num_particles = 0;
    std::vector<particle_t> particles;
};
#endif
