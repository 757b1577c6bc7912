#include <slam/occupancy_grid.hpp>
#include <algorithm>
#include <atomic>
#include <climits>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace {
std::atomic<uint64_t> g_nextGeneration(1);     // process-wide: no two grid states ever share a generation
const std::size_t kHistory = 32;
}

OccupancyGrid::OccupancyGrid(void)
: width_(0), height_(0), metersPerCell_(0.05f), cellsPerMeter_(1.0 / metersPerCell_), globalOrigin_(0, 0), generation_(0),
  writeSeq_(0), historyFloor_(0)
{
    wholeGridChanged();
}

OccupancyGrid::OccupancyGrid(float widthInMeters, float heightInMeters, float metersPerCell)
: metersPerCell_(metersPerCell), globalOrigin_(-widthInMeters / 2.0f, -heightInMeters / 2.0f), generation_(0),
  writeSeq_(0), historyFloor_(0)
{
    if (!(widthInMeters > 0.0f) || !(heightInMeters > 0.0f) || !(metersPerCell > 0.0f) ||
        metersPerCell > widthInMeters || metersPerCell > heightInMeters)
        throw std::invalid_argument("OccupancyGrid: extents and resolution must be positive, resolution <= extents");
    cellsPerMeter_ = 1.0f / metersPerCell_;
    width_ = widthInMeters * cellsPerMeter_;     // float product truncated, like the reference's constructor
    height_ = heightInMeters * cellsPerMeter_;
    cells_.assign(static_cast<std::size_t>(width_) * height_, 0);
    wholeGridChanged();
}

OccupancyGrid::OccupancyGrid(const OccupancyGrid& o)
: cells_(o.cells_), width_(o.width_), height_(o.height_), metersPerCell_(o.metersPerCell_), cellsPerMeter_(o.cellsPerMeter_),
  globalOrigin_(o.globalOrigin_), generation_(0), writeSeq_(0), historyFloor_(0)
{
    wholeGridChanged();
}

OccupancyGrid& OccupancyGrid::operator=(const OccupancyGrid& o)
{
    if (this != &o) {
        cells_ = o.cells_;
        width_ = o.width_; height_ = o.height_;
        metersPerCell_ = o.metersPerCell_; cellsPerMeter_ = o.cellsPerMeter_;
        globalOrigin_ = o.globalOrigin_;
        wholeGridChanged();
    }
    return *this;
}

void OccupancyGrid::wholeGridChanged(void)
{
    generation_ = g_nextGeneration.fetch_add(1);
    open_.x0 = open_.y0 = INT_MAX;
    open_.x1 = open_.y1 = INT_MIN;
    open_.firstSeq = open_.lastSeq = writeSeq_;
    history_.clear();
    historyFloor_ = writeSeq_;
}

void OccupancyGrid::setOrigin(float x, float y)
{
    reset();
    globalOrigin_.x -= x;
    globalOrigin_.y -= y;
}

void OccupancyGrid::reset(void)
{
    std::fill(cells_.begin(), cells_.end(), 0);
    wholeGridChanged();
}

bool OccupancyGrid::isCellInGrid(int x, int y) const
{
    return x >= 0 && x < width_ && y >= 0 && y < height_;
}

CellOdds OccupancyGrid::logOdds(int x, int y) const
{
    return isCellInGrid(x, y) ? cells_[cellIndex(x, y)] : 0;
}

void OccupancyGrid::setLogOdds(int x, int y, CellOdds value)
{
    if (!isCellInGrid(x, y)) return;
    touch(x, y);
    cells_[cellIndex(x, y)] = value;
}

bool OccupancyGrid::changesSince(uint64_t seq, int& x0, int& y0, int& x1, int& y1, bool& needFull) const
{
    needFull = false;
    if (seq >= writeSeq_) return false;
    if (open_.x1 >= open_.x0) {            // close the rectangle being collected: it becomes one history entry
        history_.push_back(open_);
        open_.x0 = open_.y0 = INT_MAX;
        open_.x1 = open_.y1 = INT_MIN;
        if (history_.size() > kHistory) {
            historyFloor_ = history_.front().lastSeq;
            history_.erase(history_.begin());
        }
    }
    if (seq < historyFloor_) { needFull = true; return true; }
    int ax0 = INT_MAX, ay0 = INT_MAX, ax1 = INT_MIN, ay1 = INT_MIN;
    for (std::size_t i = 0; i < history_.size(); ++i) {
        const Span& sp = history_[i];
        if (sp.lastSeq <= seq) continue;
        ax0 = std::min(ax0, sp.x0); ay0 = std::min(ay0, sp.y0);
        ax1 = std::max(ax1, sp.x1); ay1 = std::max(ay1, sp.y1);
    }
    x0 = std::max(ax0, 0); y0 = std::max(ay0, 0);
    x1 = std::min(ax1, width_ - 1); y1 = std::min(ay1, height_ - 1);
    return x1 >= x0 && y1 >= y0;
}

uint64_t OccupancyGrid::noteExternalWrite(int x0, int y0, int x1, int y1)
{
    touch(x0, y0);
    touch(x1, y1);
    return writeSeq_;
}

occupancy_grid_t OccupancyGrid::toLCM(void) const
{
    occupancy_grid_t msg;
    msg.origin_x = globalOrigin_.x;
    msg.origin_y = globalOrigin_.y;
    msg.meters_per_cell = metersPerCell_;
    msg.width = width_;
    msg.height = height_;
    msg.num_cells = static_cast<int32_t>(cells_.size());
    msg.cells = cells_;
    return msg;
}

void OccupancyGrid::fromLCM(const occupancy_grid_t& msg)
{
    globalOrigin_.x = msg.origin_x;
    globalOrigin_.y = msg.origin_y;
    metersPerCell_ = msg.meters_per_cell;
    cellsPerMeter_ = 1.0f / msg.meters_per_cell;
    width_ = msg.width;
    height_ = msg.height;
    cells_ = msg.cells;
    wholeGridChanged();
}

// ASCII format of the reference's .map files: "origin_x origin_y width height meters_per_cell" then height rows of
// width integers.
bool OccupancyGrid::saveToFile(const std::string& filename) const
{
    std::ofstream out(filename);
    if (!out.is_open()) {
        std::cerr << "ERROR: OccupancyGrid::saveToFile: cannot open " << filename << '\n';
        return false;
    }
    out << globalOrigin_.x << ' ' << globalOrigin_.y << ' ' << width_ << ' ' << height_ << ' ' << metersPerCell_ << '\n';
    for (int y = 0; y < height_; ++y) {
        for (int x = 0; x < width_; ++x) out << static_cast<int>(cells_[cellIndex(x, y)]) << ' ';
        out << '\n';
    }
    return out.good();
}

// Like the reference, the file's resolution replaces metersPerCell_ but cellsPerMeter_ keeps its constructor value
// (occupancy_grid.cpp:151-159 there); the device mirror receives cellsPerMeter() explicitly, so both stay consistent.
bool OccupancyGrid::loadFromFile(const std::string& filename)
{
    std::ifstream in(filename);
    if (!in.is_open()) {
        std::cerr << "ERROR: OccupancyGrid::loadFromFile: cannot open " << filename << '\n';
        return false;
    }
    int w = -1, h = -1;
    float ox = 0, oy = 0, mpc = 0;
    in >> ox >> oy >> w >> h >> mpc;
    if (!in || w <= 0 || h <= 0 || !(mpc > 0.0f)) {
        std::cerr << "ERROR: OccupancyGrid::loadFromFile: bad header in " << filename << '\n';
        return false;
    }
    globalOrigin_.x = ox; globalOrigin_.y = oy;
    width_ = w; height_ = h; metersPerCell_ = mpc;
    cells_.assign(static_cast<std::size_t>(w) * h, 0);
    int v = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            in >> v;
            cells_[cellIndex(x, y)] = static_cast<CellOdds>(v);
        }
    wholeGridChanged();
    return true;
}
