// C entry point for replaying a recorded (or synthesised) sensor log through HeadlessSLAM; bound by tests/ with ctypes.
// The oracle harness exports ref_replay_run with the same signature, driving the reference's own classes.
#include <slam/headless_slam.hpp>
#include <slam/cuda/device_filter.hpp>
#include <common/lcm_log.hpp>
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

// The same replay fed from an LCM event log (the format `lcm-logger` writes and `lcm-logplayer` feeds to `slam`): events
// on LIDAR / ODOMETRY (mbot_channels.h:9,13) are decoded and delivered in file order, everything else is skipped.
// counts_out[0..5] = events read, scans delivered, odometry delivered, undecodable payloads, fingerprint mismatches,
// resyncs after damaged framing.
extern "C" int b200_replay_log(const char* log_path, const int8_t* cells, int w, int h, float ox, float oy, float mpc,
                               int have_map, int num_particles, int mode, int hit_odds, int miss_odds,
                               float max_laser_distance, const float* initial_pose3, unsigned rand_seed, uint64_t filter_seed,
                               int max_poses, float* poses_out, int8_t* final_map_out, int* iterations_out, int* counts_out,
                               char* err, int err_len)
{
    try {
        LcmLogReader log;
        if (!log.open(log_path)) { std::snprintf(err, err_len, "cannot open %s", log_path); return -101; }
        HeadlessSLAM slam(num_particles, static_cast<HeadlessSLAM::Mode>(mode), (int8_t)hit_odds, (int8_t)miss_odds,
                          max_laser_distance);
        if (have_map) {
            occupancy_grid_t grid;
            grid.origin_x = ox; grid.origin_y = oy; grid.meters_per_cell = mpc; grid.width = w; grid.height = h;
            grid.num_cells = w * h;
            grid.cells.assign(cells, cells + (size_t)w * h);
            slam.setMap(grid);
        }
        pose_xyt_t init;
        init.x = initial_pose3[0]; init.y = initial_pose3[1]; init.theta = initial_pose3[2];
        slam.setInitialPose(init);
        slam.filter().setSeed(filter_seed);
        srand(rand_seed);
        const int64_t fpLidar = lidarFingerprint(), fpOdom = odometryFingerprint();
        int counts[6] = {0, 0, 0, 0, 0, 0};
        int iter = 0;
        LcmLogEvent ev;
        while (log.next(ev)) {
            ++counts[0];
            int64_t fp = 0;
            if (ev.channel == "LIDAR") {
                lidar_t s;
                if (!decodeLidar(ev.data, s, &fp)) { ++counts[3]; continue; }
                if (fp != fpLidar) ++counts[4];
                slam.handleLaser(s);
                ++counts[1];
            } else if (ev.channel == "ODOMETRY") {
                pose_xyt_t o;
                if (!decodeOdometry(ev.data, o, &fp)) { ++counts[3]; continue; }
                if (fp != fpOdom) ++counts[4];
                slam.handleOdometry(o);
                ++counts[2];
            } else {
                continue;
            }
            while (slam.isReadyToUpdate()) {
                const bool ok = slam.runSLAMIteration();
                const pose_xyt_t& p = slam.currentPose();
                if (iter < max_poses) {
                    poses_out[5 * iter + 0] = p.x; poses_out[5 * iter + 1] = p.y; poses_out[5 * iter + 2] = p.theta;
                    poses_out[5 * iter + 3] = ok ? 1.0f : 0.0f;
                    poses_out[5 * iter + 4] = (float)(p.utime % 1000000000LL) * 1e-6f;
                }
                ++iter;
            }
        }
        counts[5] = log.resyncs();
        *iterations_out = iter;
        for (int k = 0; k < 6; ++k) counts_out[k] = counts[k];
        if (final_map_out)
            for (int y = 0; y < slam.map().heightInCells() && y < h; ++y)
                for (int x = 0; x < slam.map().widthInCells() && x < w; ++x)
                    final_map_out[(size_t)y * w + x] = slam.map().logOdds(x, y);
        return 0;
    } catch (const b200::EngineError& e) {
        std::snprintf(err, err_len, "engine error %d: %s", e.code(), e.what());
        return e.code();
    } catch (const std::exception& e) {
        std::snprintf(err, err_len, "%s", e.what());
        return -100;
    }
}

// Decodes one log without touching the GPU (CPU tests): events, scans, odometry messages, undecodable payloads,
// fingerprint mismatches, resyncs; first_scan_out[0..3] = num_ranges, ranges[0], thetas[1], (float)(times[2] % 1e6) of
// the first scan; fingerprints_out[0..1] = the fingerprints this build expects for lidar_t and odometry_t.
extern "C" int b200_scan_log(const char* log_path, int* counts_out, float* first_scan_out, int64_t* fingerprints_out)
{
    LcmLogReader log;
    if (!log.open(log_path)) return -101;
    const int64_t fpLidar = lidarFingerprint(), fpOdom = odometryFingerprint();
    fingerprints_out[0] = fpLidar; fingerprints_out[1] = fpOdom;
    int counts[6] = {0, 0, 0, 0, 0, 0};
    LcmLogEvent ev;
    while (log.next(ev)) {
        ++counts[0];
        int64_t fp = 0;
        if (ev.channel == "LIDAR") {
            lidar_t s;
            if (!decodeLidar(ev.data, s, &fp)) { ++counts[3]; continue; }
            if (fp != fpLidar) ++counts[4];
            if (counts[1] == 0 && s.num_ranges >= 3) {
                first_scan_out[0] = (float)s.num_ranges; first_scan_out[1] = s.ranges[0]; first_scan_out[2] = s.thetas[1];
                first_scan_out[3] = (float)(s.times[2] % 1000000);
            }
            ++counts[1];
        } else if (ev.channel == "ODOMETRY") {
            pose_xyt_t o;
            if (!decodeOdometry(ev.data, o, &fp)) { ++counts[3]; continue; }
            if (fp != fpOdom) ++counts[4];
            ++counts[2];
        }
    }
    counts[5] = log.resyncs();
    for (int k = 0; k < 6; ++k) counts_out[k] = counts[k];
    return 0;
}
