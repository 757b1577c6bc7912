// Mapping -- occupancy-grid update from one scan: +hit at ray endpoints, -miss along the rays, saturating int8.
// Same interface as the reference's src/slam/mapping.hpp:25-34.  By default it runs on the host, as in the reference: it
// writes the map the particle filter reads, and OccupancyGrid's dirty rectangle carries those writes to the device
// mirror.  useDeviceMirror() moves the update onto the filter's map mirror instead (mcl_map_update, bit-identical) and
// reads the changed rectangle back so the host grid stays the authoritative copy for saving / publishing.
#ifndef B200_SLAM_MAPPING_HPP
#define B200_SLAM_MAPPING_HPP

#include <lcmtypes/pose_xyt_t.hpp>
#include <cstdint>

class OccupancyGrid;
class ParticleFilter;
class lidar_t;
struct adjusted_ray_t;

class Mapping
{
public:
    Mapping(float maxLaserDistance, int8_t hitOdds, int8_t missOdds);
    void updateMap(const lidar_t& scan, const pose_xyt_t& pose, OccupancyGrid& map);

    // ---- extension (not in the reference) ----
    /// Run updateMap on `filter`'s device mirror of the map (nullptr: back to the host loops).
    void useDeviceMirror(ParticleFilter* filter) { deviceFilter_ = filter; }

private:
    const float kMaxLaserDistance_;
    const int8_t kHitOdds_;
    const int8_t kMissOdds_;
    pose_xyt_t previousPose_;
    bool initialized_;
    ParticleFilter* deviceFilter_ = nullptr;

    void endpointCell(const adjusted_ray_t& ray, const OccupancyGrid& map, float& startX, float& startY, int& cellX,
                      int& cellY) const;
    void raiseOdds(int x, int y, OccupancyGrid& map);
    void lowerOdds(int x, int y, OccupancyGrid& map);
    void clearAlongRay(int x1, int y1, int x2, int y2, OccupancyGrid& map);
    void updateMapOnDevice(const lidar_t& scan, const pose_xyt_t& pose, OccupancyGrid& map);
};

#endif
