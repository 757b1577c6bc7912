#include <slam/cuda/device_filter.hpp>
#include <cstdlib>
#include <vector>

namespace b200 {

int defaultDevice(void)
{
    const char* s = std::getenv("B200_MCL_DEVICE");
    return s ? std::atoi(s) : 0;
}

bool defaultLegacyEqualUtime(void)
{
    const char* s = std::getenv("B200_MCL_LEGACY_UTIME");
    return s && std::atoi(s) != 0;
}

DeviceFilter::DeviceFilter(int64_t numParticles, int device, const mcl_params* params)
: engine_(nullptr), numParticles_(numParticles), mirroredGeneration_(0), mirroredSeq_(0)
{
    mcl_params defaults;
    if (!params) {
        mcl_default_params(&defaults);
        defaults.legacy_equal_utime = defaultLegacyEqualUtime() ? 1 : 0;
        params = &defaults;
    }
    int rc = mcl_create(params, numParticles, device, &engine_);
    if (rc != MCL_OK) throw EngineError(rc, std::string("mcl_create: ") + mcl_last_error(nullptr));
}

DeviceFilter::~DeviceFilter(void)
{
    mcl_destroy(engine_);
}

void DeviceFilter::check(int rc) const
{
    if (rc != MCL_OK) throw EngineError(rc, mcl_last_error(engine_));
}

void DeviceFilter::syncMap(const OccupancyGrid& map)
{
    // generations are process-wide unique, so an equal generation means: the grid state this mirror was filled from,
    // plus cell writes -- and those carry sequence numbers, of which this mirror remembers the last one it has seen
    bool full = mirroredGeneration_ != map.generation();
    int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    if (!full) {
        bool needFull = false;
        if (!map.changesSince(mirroredSeq_, x0, y0, x1, y1, needFull)) { mirroredSeq_ = map.writeSeq(); return; }
        full = needFull;
    }
    if (full) {
        check(mcl_set_map(engine_, map.data(), map.widthInCells(), map.heightInCells(), map.originInGlobalFrame().x,
                          map.originInGlobalFrame().y, map.metersPerCell(), map.cellsPerMeter()));
        mirroredGeneration_ = map.generation();
    } else {
        check(mcl_update_map_rect(engine_, x0, y0, x1 - x0 + 1, y1 - y0 + 1,
                                  map.data() + static_cast<std::size_t>(y0) * map.widthInCells() + x0,
                                  map.widthInCells()));
    }
    mirroredSeq_ = map.writeSeq();
}

void DeviceFilter::noteMirrorIsAheadOf(const OccupancyGrid& map, uint64_t seq)
{
    if (mirroredGeneration_ == map.generation() && seq == mirroredSeq_ + 2) mirroredSeq_ = seq;
}

}  // namespace b200
