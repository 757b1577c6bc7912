#include <slam/cuda/device_filter.hpp>
#include <cstdlib>
#include <vector>

namespace b200 {

int defaultDevice(void)
{
    const char* s = std::getenv("B200_MCL_DEVICE");
    return s ? std::atoi(s) : 0;
}

DeviceFilter::DeviceFilter(int64_t numParticles, int device, const mcl_params* params)
: engine_(nullptr), numParticles_(numParticles), mirrored_(nullptr), mirroredGeneration_(0)
{
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
    if (mirrored_ != &map || mirroredGeneration_ != map.generation()) {
        check(mcl_set_map(engine_, map.data(), map.widthInCells(), map.heightInCells(), map.originInGlobalFrame().x,
                          map.originInGlobalFrame().y, map.metersPerCell(), map.cellsPerMeter()));
        mirrored_ = &map;
        mirroredGeneration_ = map.generation();
        map.clearDirty();
        return;
    }
    int x0, y0, x1, y1;
    if (map.dirtyRect(x0, y0, x1, y1)) {
        check(mcl_update_map_rect(engine_, x0, y0, x1 - x0 + 1, y1 - y0 + 1,
                                  map.data() + static_cast<std::size_t>(y0) * map.widthInCells() + x0,
                                  map.widthInCells()));
        map.clearDirty();
    }
}

}  // namespace b200
