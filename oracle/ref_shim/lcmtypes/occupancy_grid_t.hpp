// Oracle-only stand-in for lcmtypes/occupancy_grid_t.lcm:1-14.
#ifndef ORACLE_SHIM_OCCUPANCY_GRID_T_HPP
#define ORACLE_SHIM_OCCUPANCY_GRID_T_HPP
#include <cstdint>
#include <vector>
class occupancy_grid_t
{
public:
    int64_t utime;
    float origin_x;
    float origin_y;
    float meters_per_cell;
    int32_t width;
    int32_t height;
    int32_t num_cells;
    std::vector<int8_t> cells;
};
#endif
