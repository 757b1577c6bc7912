// ActionModel -- odometry rotate-translate-rotate motion model with the reference's public interface
// (src/slam/action_model.hpp:35-54).  updateAction is the scalar host half (once per update); the per-particle half
// runs on the GPU inside ParticleFilter::updateFilter, drawing its noise from a counter-based Philox4x32-10 stream.
// applyAction(sample) keeps the one-particle call for API compatibility by running the same kernel on a tiny batch.
#ifndef B200_SLAM_ACTION_MODEL_HPP
#define B200_SLAM_ACTION_MODEL_HPP

#include <lcmtypes/pose_xyt_t.hpp>
#include <slam/cuda/mcl_cuda.h>
#include <memory>

class particle_t;
namespace b200 { class DeviceFilter; }

class ActionModel
{
public:
    ActionModel(void);
    ~ActionModel(void);

    /// Latches the odometry delta since the previous call.  \return whether the robot moved (action_model.cpp:52)
    bool updateAction(const pose_xyt_t& odometry);

    /// Moves one sample by a noisy copy of the latched action; parent_pose becomes the sample's old pose.
    particle_t applyAction(const particle_t& sample);

    /// The latched action, for ParticleFilter to hand to the engine.
    const mcl_action_t& action(void) const { return action_; }
    int64_t utime(void) const { return utime_; }

private:
    mcl_action_t action_;
    int64_t utime_;                                   // utime given to moved particles: the latched odometry's
    std::unique_ptr<b200::DeviceFilter> single_;      // lazily created 2-particle engine behind applyAction
    uint64_t singleCalls_;
};

#endif
