// SensorModel -- scan score of a particle against the int8 log-odds grid, reference interface
// (src/slam/sensor_model.hpp:28-38).  The score is the reference's: per ray, the endpoint cell's log-odds if positive,
// else half the log-odds of the cell one Bresenham step toward / away from the robot; summed over rays.  The
// arithmetic runs in the CUDA engine (csrc/mcl_device.cuh); likelihood() of one particle is a batch of one.
#ifndef B200_SLAM_SENSOR_MODEL_HPP
#define B200_SLAM_SENSOR_MODEL_HPP

#include <memory>
#include <vector>

class lidar_t;
class OccupancyGrid;
class particle_t;
namespace b200 { class DeviceFilter; }

class SensorModel
{
public:
    SensorModel(void);
    ~SensorModel(void);

    double likelihood(const particle_t& particle, const lidar_t& scan, const OccupancyGrid& map);

    /// Batch form: scores[i] for particles[i]; all particles must share pose.utime and parent_pose.utime.
    std::vector<double> likelihoods(const std::vector<particle_t>& particles, const lidar_t& scan,
                                    const OccupancyGrid& map);

private:
    std::unique_ptr<b200::DeviceFilter> device_;
};

#endif
