#include <slam/sensor_model.hpp>
#include <slam/cuda/device_filter.hpp>
#include <slam/occupancy_grid.hpp>
#include <lcmtypes/lidar_t.hpp>
#include <lcmtypes/particle_t.hpp>

SensorModel::SensorModel(void) {}
SensorModel::~SensorModel(void) = default;

std::vector<double> SensorModel::likelihoods(const std::vector<particle_t>& particles, const lidar_t& scan,
                                             const OccupancyGrid& map)
{
    std::vector<particle_t> batch(particles);
    if (batch.empty()) return std::vector<double>();
    if (batch.size() < 2) batch.push_back(batch.front());          // the engine holds at least two particles
    const int64_t n = static_cast<int64_t>(batch.size());
    if (!device_ || device_->numParticles() != n) device_.reset(new b200::DeviceFilter(n, b200::defaultDevice()));
    device_->syncMap(map);
    device_->check(mcl_import_particles(device_->engine(), reinterpret_cast<const mcl_particle_t*>(batch.data()), n));
    std::vector<double> scores(batch.size());
    device_->check(mcl_score(device_->engine(), scan.ranges.data(), scan.thetas.data(), scan.times.data(),
                             scan.num_ranges, scores.data()));
    scores.resize(particles.size());
    return scores;
}

double SensorModel::likelihood(const particle_t& particle, const lidar_t& scan, const OccupancyGrid& map)
{
    return likelihoods(std::vector<particle_t>(1, particle), scan, map).front();
}
