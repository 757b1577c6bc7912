#include <slam/particle_filter.hpp>
#include <slam/cuda/device_filter.hpp>
#include <slam/occupancy_grid.hpp>
#include <lcmtypes/lidar_t.hpp>
#include <cstdlib>
#include <stdexcept>

ParticleFilter::ParticleFilter(int numParticles)
: kNumParticles_(numParticles), seed_(0x6d636cULL), maxExported_(numParticles), exportWeighted_(false), injectedNoise_(nullptr)
{
    if (numParticles <= 1) throw std::invalid_argument("ParticleFilter needs more than one particle");
    device_.reset(new b200::DeviceFilter(numParticles, b200::defaultDevice()));
}

ParticleFilter::ParticleFilter(int numParticles, const mcl_params& params)
: kNumParticles_(numParticles), seed_(0x6d636cULL), maxExported_(numParticles), exportWeighted_(false), injectedNoise_(nullptr)
{
    if (numParticles <= 1) throw std::invalid_argument("ParticleFilter needs more than one particle");
    device_.reset(new b200::DeviceFilter(numParticles, b200::defaultDevice(), &params));
}

ParticleFilter::~ParticleFilter(void) = default;

void ParticleFilter::setSeed(uint64_t seed)
{
    seed_ = seed;
    srand(static_cast<unsigned>(seed));
}

void ParticleFilter::initializeFilterAtPose(const pose_xyt_t& pose)
{
    posteriorPose_ = pose;
    device_->check(mcl_init_at_pose(device_->engine(), pose.x, pose.y, pose.theta, pose.utime, seed_));
}

pose_xyt_t ParticleFilter::updateFilter(const pose_xyt_t& odometry, const lidar_t& laser, const OccupancyGrid& map)
{
    const bool moved = actionModel_.updateAction(odometry);
    if (moved) {
        device_->syncMap(map);
        // the systematic-resampling offset comes from libc rand() exactly like the reference
        // (particle_filter.cpp:89-92 there): r = rand()/RAND_MAX * 1/N
        const double r = (static_cast<double>(rand()) / static_cast<double>(RAND_MAX)) * (1.0 / kNumParticles_);
        mcl_pose_t est;
        device_->check(mcl_update(device_->engine(), &actionModel_.action(), odometry.utime, laser.ranges.data(),
                                  laser.thetas.data(), laser.times.data(), laser.num_ranges, r, injectedNoise_, &est));
        injectedNoise_ = nullptr;
        posteriorPose_.x = est.x;
        posteriorPose_.y = est.y;
        posteriorPose_.theta = est.theta;
    }
    posteriorPose_.utime = odometry.utime;
    return posteriorPose_;
}

pose_xyt_t ParticleFilter::updateFilterActionOnly(const pose_xyt_t& odometry)
{
    if (actionModel_.updateAction(odometry))
    {
        device_->check(mcl_update_action_only(device_->engine(), &actionModel_.action(), odometry.utime,
                                              injectedNoise_));
        injectedNoise_ = nullptr;
    }
    posteriorPose_ = odometry;
    return posteriorPose_;
}

pose_xyt_t ParticleFilter::poseEstimate(void) const
{
    return posteriorPose_;
}

particles_t ParticleFilter::particles(void) const
{
    particles_t out;
    const int64_t cap = maxExported_ > 0 ? maxExported_ : kNumParticles_;
    if (exportWeighted_ && cap < kNumParticles_) {
        // a weighted draw of `cap` particles (equal weights) instead of every k-th one: what a viewer of a large cloud
        // wants to see is where the probability mass is
        out.particles.resize(static_cast<std::size_t>(cap));
        int64_t n = 0;
        device_->check(mcl_export_weighted(device_->engine(), reinterpret_cast<mcl_particle_t*>(out.particles.data()), cap,
                                           0.5, &n));
        out.particles.resize(static_cast<std::size_t>(n));
        out.num_particles = static_cast<int32_t>(n);
        out.utime = posteriorPose_.utime;
        return out;
    }
    const int64_t stride = (kNumParticles_ + cap - 1) / cap;
    out.particles.resize(static_cast<std::size_t>((kNumParticles_ + stride - 1) / stride));
    int64_t n = 0;
    device_->check(mcl_export_particles(device_->engine(), reinterpret_cast<mcl_particle_t*>(out.particles.data()),
                                        static_cast<int64_t>(out.particles.size()), stride, &n));
    out.particles.resize(static_cast<std::size_t>(n));
    out.num_particles = static_cast<int32_t>(n);
    out.utime = posteriorPose_.utime;
    return out;
}

void ParticleFilter::setParticles(const particles_t& cloud)
{
    if (static_cast<int>(cloud.particles.size()) != kNumParticles_)
        throw std::invalid_argument("setParticles needs exactly numParticles particles");
    device_->check(mcl_import_particles(device_->engine(), reinterpret_cast<const mcl_particle_t*>(cloud.particles.data()),
                                        kNumParticles_));
}

mcl_stats ParticleFilter::stats(void) const
{
    mcl_stats s;
    device_->check(mcl_get_stats(device_->engine(), &s));
    return s;
}
