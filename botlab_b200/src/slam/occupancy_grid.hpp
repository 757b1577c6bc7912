// OccupancyGrid -- host-side int8 log-odds grid with the reference's public interface (src/slam/occupancy_grid.hpp:
// 62-194 in the reference) plus what the device mirror needs: a dirty rectangle of cells written since the mirror was
// last refreshed.  The map changes on every SLAM iteration, even with --localization-only (the mode test at
// slam.cpp:276 is always true), so ParticleFilter patches the mirror with mcl_update_map_rect before each update.
#ifndef B200_SLAM_OCCUPANCY_GRID_HPP
#define B200_SLAM_OCCUPANCY_GRID_HPP

#include <common/point.hpp>
#include <lcmtypes/occupancy_grid_t.hpp>
#include <cstdint>
#include <string>
#include <vector>

typedef int8_t CellOdds;   ///< log-odds of occupancy: > 0 occupied, < 0 free, 0 unknown

class OccupancyGrid
{
public:
    /// Empty grid, 0.05 m cells, origin (0, 0).
    OccupancyGrid(void);
    /// Grid centred on the global origin.  \pre all three > 0 and metersPerCell <= both extents
    OccupancyGrid(float widthInMeters, float heightInMeters, float metersPerCell);

    int   widthInCells(void) const { return width_; }
    float widthInMeters(void) const { return width_ * metersPerCell_; }
    int   heightInCells(void) const { return height_; }
    float heightInMeters(void) const { return height_ * metersPerCell_; }
    float metersPerCell(void) const { return metersPerCell_; }
    float cellsPerMeter(void) const { return cellsPerMeter_; }
    Point<float> originInGlobalFrame(void) const { return globalOrigin_; }

    void setOrigin(float x, float y);
    void reset(void);

    bool isCellInGrid(int x, int y) const;
    CellOdds logOdds(int x, int y) const;              ///< 0 outside the grid
    void setLogOdds(int x, int y, CellOdds logOdds);    ///< ignored outside the grid

    /// Unchecked access.  The non-const form marks the cell dirty (the caller may write through the reference).
    CellOdds& operator()(int x, int y) { touch(x, y); return cells_[cellIndex(x, y)]; }
    CellOdds  operator()(int x, int y) const { return cells_[cellIndex(x, y)]; }

    occupancy_grid_t toLCM(void) const;
    void fromLCM(const occupancy_grid_t& gridMessage);
    bool saveToFile(const std::string& filename) const;
    bool loadFromFile(const std::string& filename);

    // ---- device-mirror support (not in the reference) ----
    const CellOdds* data(void) const { return cells_.data(); }
    /// Writable storage for values read back FROM the device mirror: writing through it does not mark cells dirty.
    CellOdds* mirrorData(void) { return cells_.data(); }
    /// Bumped whenever the geometry or the whole content changes (ctor, reset, setOrigin, fromLCM, loadFromFile).
    uint64_t generation(void) const { return generation_; }
    /// Bounding box [x0,x1] x [y0,y1] of cells written since clearDirty(); false if none.
    bool dirtyRect(int& x0, int& y0, int& x1, int& y1) const;
    void clearDirty(void) const;

private:
    std::vector<CellOdds> cells_;
    int width_;
    int height_;
    float metersPerCell_;
    float cellsPerMeter_;
    Point<float> globalOrigin_;

    uint64_t generation_;
    mutable int dirtyX0_, dirtyY0_, dirtyX1_, dirtyY1_;

    int cellIndex(int x, int y) const { return y * width_ + x; }
    void touch(int x, int y)
    {
        if (x < dirtyX0_) dirtyX0_ = x;
        if (x > dirtyX1_) dirtyX1_ = x;
        if (y < dirtyY0_) dirtyY0_ = y;
        if (y > dirtyY1_) dirtyY1_ = y;
    }
    void wholeGridChanged(void);
};

#endif  // B200_SLAM_OCCUPANCY_GRID_HPP
