#include <slam/mapping.hpp>
#include <slam/moving_laser_scan.hpp>
#include <slam/occupancy_grid.hpp>
#include <slam/particle_filter.hpp>
#include <slam/cuda/device_filter.hpp>
#include <lcmtypes/lidar_t.hpp>
#include <cmath>
#include <cstdlib>
#include <limits>

Mapping::Mapping(float maxLaserDistance, int8_t hitOdds, int8_t missOdds)
: kMaxLaserDistance_(maxLaserDistance), kHitOdds_(hitOdds), kMissOdds_(missOdds), initialized_(false)
{
}

// Until the first call has latched a previous pose the reference changes no cell (mapping.cpp:73-76,87-90 there).
void Mapping::updateMap(const lidar_t& scan, const pose_xyt_t& pose, OccupancyGrid& map)
{
    if (!initialized_) previousPose_ = pose;
    if (deviceFilter_) {
        updateMapOnDevice(scan, pose, map);
        initialized_ = true;
        previousPose_ = pose;
        return;
    }
    const MovingLaserScan rays(scan, previousPose_, pose);
    for (const adjusted_ray_t& ray : rays) {           // occupied endpoints first ...
        if (!(ray.range <= kMaxLaserDistance_)) continue;
        float sx, sy; int cx, cy;
        endpointCell(ray, map, sx, sy, cx, cy);
        if (map.isCellInGrid(cx, cy)) raiseOdds(cx, cy, map);
    }
    for (const adjusted_ray_t& ray : rays) {           // ... then the free space the rays crossed
        if (!(ray.range <= kMaxLaserDistance_)) continue;
        float sx, sy; int cx, cy;
        endpointCell(ray, map, sx, sy, cx, cy);
        clearAlongRay(static_cast<int>(sx), static_cast<int>(sy), cx, cy, map);
    }
    initialized_ = true;
    previousPose_ = pose;
}

// Grid position of the ray origin (double math stored as float) and the truncated endpoint cell, float arithmetic in
// the reference's order: (range * cos) * cellsPerMeter + start.
void Mapping::endpointCell(const adjusted_ray_t& ray, const OccupancyGrid& map, float& startX, float& startY,
                           int& cellX, int& cellY) const
{
    const float cpm = map.cellsPerMeter();
    startX = static_cast<float>((static_cast<double>(ray.origin.x) - map.originInGlobalFrame().x) * cpm);
    startY = static_cast<float>((static_cast<double>(ray.origin.y) - map.originInGlobalFrame().y) * cpm);
    cellX = static_cast<int>((ray.range * std::cos(ray.theta) * cpm) + startX);
    cellY = static_cast<int>((ray.range * std::sin(ray.theta) * cpm) + startY);
}

void Mapping::raiseOdds(int x, int y, OccupancyGrid& map)
{
    if (!initialized_) return;
    const int cur = map.logOdds(x, y);
    const int top = std::numeric_limits<CellOdds>::max();
    map.setLogOdds(x, y, static_cast<CellOdds>(top - cur > kHitOdds_ ? cur + kHitOdds_ : top));
}

void Mapping::lowerOdds(int x, int y, OccupancyGrid& map)
{
    if (!initialized_) return;
    const int cur = map.logOdds(x, y);
    const int bottom = std::numeric_limits<CellOdds>::min();
    map.setLogOdds(x, y, static_cast<CellOdds>(cur - kMissOdds_ > bottom ? cur - kMissOdds_ : bottom));
}

// Bresenham walk from the ray's start cell up to (not including) its endpoint cell.
void Mapping::clearAlongRay(int x1, int y1, int x2, int y2, OccupancyGrid& map)
{
    const int dx = std::abs(x2 - x1), dy = std::abs(y2 - y1);
    const int sx = x1 < x2 ? 1 : -1, sy = y1 < y2 ? 1 : -1;
    int err = dx - dy, x = x1, y = y1;
    while (x != x2 || y != y2) {
        if (map.isCellInGrid(x, y)) lowerOdds(x, y, map);
        const float e2 = 2 * err;
        if (e2 >= -dy) { err -= dy; x += sx; }
        if (e2 <= dx) { err += dx; y += sy; }
    }
}

// The same update on the filter's map mirror: bring the mirror up to date with any host-side writes, run
// mcl_map_update, read the rectangle that may have changed back into the host grid and register it as written (any other
// mirror of the grid picks it up; this one is told it already holds those values).
void Mapping::updateMapOnDevice(const lidar_t& scan, const pose_xyt_t& pose, OccupancyGrid& map)
{
    b200::DeviceFilter& dev = deviceFilter_->device();
    dev.syncMap(map);
    mcl_pose_t prev, cur;
    prev.utime = previousPose_.utime; prev.x = previousPose_.x; prev.y = previousPose_.y; prev.theta = previousPose_.theta;
    cur.utime = pose.utime; cur.x = pose.x; cur.y = pose.y; cur.theta = pose.theta;
    int rect[4] = {0, 0, 0, 0};
    dev.check(mcl_map_update(dev.engine(), &prev, &cur, initialized_ ? 1 : 0, scan.ranges.data(), scan.thetas.data(),
                             scan.times.data(), scan.num_ranges, kMaxLaserDistance_, kHitOdds_, kMissOdds_, rect));
    if (rect[2] > 0 && rect[3] > 0) {
        dev.check(mcl_read_map_rect(dev.engine(), rect[0], rect[1], rect[2], rect[3],
                                    map.mirrorData() + static_cast<std::size_t>(rect[1]) * map.widthInCells() + rect[0],
                                    map.widthInCells()));
        dev.noteMirrorIsAheadOf(map, map.noteExternalWrite(rect[0], rect[1], rect[0] + rect[2] - 1, rect[1] + rect[3] - 1));
    }
}
