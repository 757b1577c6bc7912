#include <slam/action_model.hpp>
#include <slam/cuda/device_filter.hpp>
#include <lcmtypes/particle_t.hpp>
#include <cstring>

static_assert(sizeof(pose_xyt_t) == sizeof(mcl_pose_t), "pose_xyt_t must match the C ABI layout");
static_assert(sizeof(particle_t) == sizeof(mcl_particle_t), "particle_t must match the C ABI layout");

ActionModel::ActionModel(void) : utime_(0), singleCalls_(0)
{
    mcl_action_reset(&action_);
}

ActionModel::~ActionModel(void) = default;

bool ActionModel::updateAction(const pose_xyt_t& odometry)
{
    mcl_pose_t o;
    o.utime = odometry.utime; o.x = odometry.x; o.y = odometry.y; o.theta = odometry.theta;
    utime_ = odometry.utime;    // the reference never assigns its utime_ (action_model.hpp:72); this is the intent
    return mcl_action_update(&action_, &o) != 0;
}

particle_t ActionModel::applyAction(const particle_t& sample)
{
    if (!single_) single_.reset(new b200::DeviceFilter(2, b200::defaultDevice()));
    mcl_particle_t batch[2];
    std::memcpy(&batch[0], &sample, sizeof(mcl_particle_t));
    batch[1] = batch[0];
    single_->check(mcl_import_particles(single_->engine(), batch, 2));
    // a fresh Philox stream position per call: re-seeding through init is avoided by advancing the update counter
    single_->check(mcl_apply_action(single_->engine(), &action_, utime_, nullptr));
    ++singleCalls_;
    int64_t n = 0;
    single_->check(mcl_export_particles(single_->engine(), batch, 1, 1, &n));
    particle_t out;
    std::memcpy(static_cast<void*>(&out), &batch[0], sizeof(mcl_particle_t));
    return out;
}
