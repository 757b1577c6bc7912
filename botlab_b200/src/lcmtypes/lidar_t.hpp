// Stand-in for lcm-gen's lidar_t (lcmtypes/lidar_t.lcm:1-14).
#ifndef B200_LCMTYPES_LIDAR_T_HPP
#define B200_LCMTYPES_LIDAR_T_HPP
#include <cstdint>
#include <vector>
class lidar_t
{
public:
    int64_t utime = 0;
    int32_t num_ranges = 0;
    std::vector<float> ranges;
    std::vector<float> thetas;
    std::vector<int64_t> times;
    std::vector<float> intensities;
};
#endif
