// OccupancyGrid -- host-side int8 log-odds grid with the reference's public interface (src/slam/occupancy_grid.hpp:
// 62-194 in the reference) plus what device mirrors need: a process-wide unique generation (whole-grid changes) and a
// short history of written rectangles stamped with a write sequence, so that ANY number of mirrors can each ask "what
// changed since I last looked".  The map changes on every SLAM iteration, even with --localization-only (the mode test at
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
    /// Copies get a generation of their own: a mirror of the source is never mistaken for a mirror of the copy.
    OccupancyGrid(const OccupancyGrid& other);
    OccupancyGrid& operator=(const OccupancyGrid& other);

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
    /// Drawn from a process-wide counter whenever the geometry or the whole content changes (ctors, copies, reset,
    /// setOrigin, fromLCM, loadFromFile): equal generations mean the same grid object in the same whole-grid state.
    uint64_t generation(void) const { return generation_; }
    /// Sequence number of the last cell write (setLogOdds, non-const operator(), noteExternalWrite).
    uint64_t writeSeq(void) const { return writeSeq_; }
    /// Bounding box [x0,x1] x [y0,y1] of the cells written after write sequence `seq`.  Returns false when nothing was.
    /// needFull is set when the history no longer reaches back to `seq` (the caller must take the whole grid).
    bool changesSince(uint64_t seq, int& x0, int& y0, int& x1, int& y1, bool& needFull) const;
    /// Writable storage for values read back FROM a device mirror, followed by noteExternalWrite(rect): the cells count
    /// as written (other mirrors will pick them up); the returned sequence is what the originating mirror has seen.
    CellOdds* mirrorData(void) { return cells_.data(); }
    uint64_t noteExternalWrite(int x0, int y0, int x1, int y1);

private:
    std::vector<CellOdds> cells_;
    int width_;
    int height_;
    float metersPerCell_;
    float cellsPerMeter_;
    Point<float> globalOrigin_;

    uint64_t generation_;
    uint64_t writeSeq_;
    // rectangle of the writes since the last changesSince() call, and a short history of the closed ones
    struct Span { uint64_t firstSeq, lastSeq; int x0, y0, x1, y1; };
    mutable Span open_;
    mutable std::vector<Span> history_;
    mutable uint64_t historyFloor_;     // writes with a sequence <= this are no longer in history_

    int cellIndex(int x, int y) const { return y * width_ + x; }
    void touch(int x, int y)
    {
        ++writeSeq_;
        if (open_.x1 < open_.x0) open_.firstSeq = writeSeq_;
        open_.lastSeq = writeSeq_;
        if (x < open_.x0) open_.x0 = x;
        if (x > open_.x1) open_.x1 = x;
        if (y < open_.y0) open_.y0 = y;
        if (y > open_.y1) open_.y1 = y;
    }
    void wholeGridChanged(void);
};

#endif  // B200_SLAM_OCCUPANCY_GRID_HPP
