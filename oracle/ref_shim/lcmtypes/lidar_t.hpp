// Oracle-only stand-in for lcmtypes/lidar_t.lcm:1-14.
#ifndef ORACLE_SHIM_LIDAR_T_HPP
#define ORACLE_SHIM_LIDAR_T_HPP
#include <cstdint>
#include <vector>
class lidar_t
{
public:
    int64_t utime;
    int32_t num_ranges;
    std::vector<float> ranges;
    std::vector<float> thetas;
    std::vector<int64_t> times;
    std::vector<float> intensities;
};
#endif
