// Stand-in for lcm-gen's occupancy_grid_t (lcmtypes/occupancy_grid_t.lcm:1-14).
#ifndef B200_LCMTYPES_OCCUPANCY_GRID_T_HPP
#define B200_LCMTYPES_OCCUPANCY_GRID_T_HPP
#include <cstdint>
#include <vector>
class occupancy_grid_t
{
public:
    int64_t utime = 0;
    float origin_x = 0.0f;
    float origin_y = 0.0f;
    float meters_per_cell = 0.0f;
    int32_t width = 0;
    int32_t height = 0;
    int32_t num_cells = 0;
    std::vector<int8_t> cells;
};
#endif
