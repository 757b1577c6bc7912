// C entry point for replaying a recorded (or synthesised) sensor log through HeadlessSLAM; bound by tests/ with ctypes.
// The oracle harness exports ref_replay_run with the same signature, driving the reference's own classes.
#include <slam/headless_slam.hpp>
#include <slam/cuda/device_filter.hpp>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct ReplayHooks {
    const mcl_particle_t* initCloud;
    int numParticles;
    const float* noise;      // [iteration][N][3]
    int maxIterations;
};

void plantCloud(HeadlessSLAM& slam, void* arg)
{
    ReplayHooks* h = static_cast<ReplayHooks*>(arg);
    if (!h->initCloud) return;
    particles_t cloud;
    cloud.num_particles = h->numParticles;
    cloud.particles.resize(h->numParticles);
    std::memcpy(static_cast<void*>(cloud.particles.data()), h->initCloud, sizeof(mcl_particle_t) * h->numParticles);
    slam.filter().setParticles(cloud);
}

void injectNoise(HeadlessSLAM& slam, int iteration, void* arg)
{
    ReplayHooks* h = static_cast<ReplayHooks*>(arg);
    if (h->noise && iteration < h->maxIterations)
        slam.filter().injectActionNoise(h->noise + static_cast<size_t>(iteration) * h->numParticles * 3);
}
}  // namespace

extern "C" int b200_replay_run(const int8_t* cells, int w, int h, float ox, float oy, float mpc, int have_map,
                               int num_particles, int mode, int hit_odds, int miss_odds, float max_laser_distance,
                               int num_scans, const int32_t* scan_offsets, const float* ranges, const float* thetas,
                               const int64_t* times, int num_odom, const int64_t* odom_utime, const float* odom_xyt,
                               const float* initial_pose3, unsigned rand_seed, const void* init_cloud, float* noise_io,
                               float* poses_out, int8_t* final_map_out, int* iterations_out, char* err, int err_len)
{
    try {
        HeadlessSLAM slam(num_particles, static_cast<HeadlessSLAM::Mode>(mode), (int8_t)hit_odds, (int8_t)miss_odds,
                          max_laser_distance);
        occupancy_grid_t grid;
        grid.origin_x = ox; grid.origin_y = oy; grid.meters_per_cell = mpc; grid.width = w; grid.height = h;
        grid.num_cells = w * h;
        grid.cells.assign(cells, cells + (size_t)w * h);
        if (have_map) slam.setMap(grid);   // else: full SLAM starts from the constructor's empty 10 m x 10 m grid
        pose_xyt_t init;
        init.x = initial_pose3[0]; init.y = initial_pose3[1]; init.theta = initial_pose3[2];
        slam.setInitialPose(init);
        // B200_DEVICE_MAPPING=1: Mapping::updateMap runs on the device mirror (tests replay both ways)
        const char* dm = std::getenv("B200_DEVICE_MAPPING");
        slam.setDeviceMapping(dm && dm[0] == '1');
        ReplayHooks hooks{static_cast<const mcl_particle_t*>(init_cloud), num_particles, noise_io, num_scans};
        slam.onFilterInitialized = plantCloud;
        slam.beforeLocalization = injectNoise;
        slam.hookArg = &hooks;
        srand(rand_seed);
        // deliver messages in arrival order: odometry at its utime, a scan when its last ray has been measured
        int io = 0, is = 0, iter = 0;
        while (io < num_odom || is < num_scans) {
            const int64_t to = io < num_odom ? odom_utime[io] : INT64_MAX;
            const int64_t ts = is < num_scans ? times[scan_offsets[is + 1] - 1] : INT64_MAX;
            if (to <= ts) {
                pose_xyt_t o;
                o.utime = odom_utime[io]; o.x = odom_xyt[3 * io]; o.y = odom_xyt[3 * io + 1]; o.theta = odom_xyt[3 * io + 2];
                slam.handleOdometry(o);
                ++io;
            } else {
                lidar_t s;
                const int a = scan_offsets[is], b = scan_offsets[is + 1];
                s.num_ranges = b - a;
                s.ranges.assign(ranges + a, ranges + b);
                s.thetas.assign(thetas + a, thetas + b);
                s.times.assign(times + a, times + b);
                s.intensities.assign(b - a, 0.0f);
                s.utime = s.times.back();
                slam.handleLaser(s);
                ++is;
            }
            while (slam.isReadyToUpdate()) {
                const bool ok = slam.runSLAMIteration();
                const pose_xyt_t& p = slam.currentPose();
                if (iter < num_scans) {
                    poses_out[5 * iter + 0] = p.x; poses_out[5 * iter + 1] = p.y; poses_out[5 * iter + 2] = p.theta;
                    poses_out[5 * iter + 3] = ok ? 1.0f : 0.0f;
                    poses_out[5 * iter + 4] = (float)(p.utime % 1000000000LL) * 1e-6f;
                }
                ++iter;
            }
        }
        *iterations_out = iter;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) final_map_out[(size_t)y * w + x] = slam.map().logOdds(x, y);
        return 0;
    } catch (const b200::EngineError& e) {
        std::snprintf(err, err_len, "engine error %d: %s", e.code(), e.what());
        return e.code();
    } catch (const std::exception& e) {
        std::snprintf(err, err_len, "%s", e.what());
        return -100;
    }
}
