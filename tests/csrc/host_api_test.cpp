// Drives the C++ host classes (botlab_b200/src/slam) the way OccupancyGridSLAM drives the reference's
// (slam.cpp:23,38,246,259-265): load the map, initialise the filter at a pose, then one updateFilter per scan, with
// the map mutated between updates.  Prints one JSON object; tests/test_host_api.py checks it.
// usage: host_api_test <file.map> <num_particles> <steps>
#include <slam/particle_filter.hpp>
#include <slam/moving_laser_scan.hpp>
#include <slam/occupancy_grid.hpp>
#include <slam/cuda/device_filter.hpp>
#include <lcmtypes/lidar_t.hpp>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static lidar_t katScan()
{
    lidar_t s;
    s.num_ranges = 360;
    for (int i = 0; i < 360; ++i) {
        s.ranges.push_back(1.0f + 0.002f * i);
        s.thetas.push_back((float)(i * 2 * M_PI / 360));
        s.times.push_back(1000000 + 277 * i);
        s.intensities.push_back(0.0f);
    }
    s.utime = s.times.back();
    return s;
}

static particle_t part(float x, float y, float th, int64_t t, float px, float py, float pth, int64_t pt)
{
    particle_t p;
    p.pose.x = x; p.pose.y = y; p.pose.theta = th; p.pose.utime = t;
    p.parent_pose.x = px; p.parent_pose.y = py; p.parent_pose.theta = pth; p.parent_pose.utime = pt;
    p.weight = 0.5;
    return p;
}

// ray-march a scan from (x, y, th) like src/sim/lidar.py: first cell > 0 or 8 m
static lidar_t marchScan(const OccupancyGrid& map, float x, float y, float th, int64_t tEnd)
{
    lidar_t s;
    s.num_ranges = 360;
    for (int i = 0; i < 360; ++i) {
        const float beam = (float)(2 * M_PI * i / 360);
        const double ang = th - beam;
        double d = 0.025;
        for (; d < 8.0; d += 0.025) {
            const int cx = (int)std::floor((x + d * std::cos(ang) - map.originInGlobalFrame().x) * map.cellsPerMeter());
            const int cy = (int)std::floor((y + d * std::sin(ang) - map.originInGlobalFrame().y) * map.cellsPerMeter());
            if (map.logOdds(cx, cy) > 0) break;
        }
        s.ranges.push_back((float)d);
        s.thetas.push_back(beam);
        s.times.push_back(tEnd - 100000 + (int64_t)(i + 1) * 100000 / 360);
        s.intensities.push_back(0.0f);
    }
    s.utime = tEnd;
    return s;
}

int main(int argc, char** argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s <file.map> <num_particles> <steps>\n", argv[0]); return 2; }
    const int n = std::atoi(argv[2]), steps = std::atoi(argv[3]);
    try {
        OccupancyGrid map(10.0f, 10.0f, 0.05f);                  // slam.cpp:23
        if (!map.loadFromFile(argv[1])) return 3;                // slam.cpp:38
        std::printf("{\"width\": %d, \"height\": %d, \"cpm\": %.9g", map.widthInCells(), map.heightInCells(),
                    map.cellsPerMeter());

        // SensorModel::likelihood known answers
        SensorModel sm;
        const lidar_t ks = katScan();
        const particle_t kp[4] = {part(0, 0, 0, 1100000, 0, 0, 0, 1100000),
                                  part(0.5f, -0.25f, 0.3f, 1100000, 0.48f, -0.26f, 0.28f, 1000000),
                                  part(-2.0f, 1.5f, -2.5f, 1100000, -2.02f, 1.49f, -2.45f, 1000000),
                                  part(1.0f, 1.0f, 3.1f, 1100000, 0.98f, 1.0f, -3.1f, 1000000)};
        std::printf(", \"kat_scores\": [");
        for (int i = 0; i < 4; ++i) std::printf("%s%.17g", i ? ", " : "", sm.likelihood(kp[i], ks, map));
        std::printf("]");

        // MovingLaserScan known answers (B5-B7)
        MovingLaserScan ms(ks, kp[1].parent_pose, kp[1].pose);
        std::printf(", \"ray0\": [%.9g, %.9g, %.9g, %.9g], \"ray200\": [%.9g, %.9g, %.9g, %.9g], \"rays\": %zu",
                    ms[0].origin.x, ms[0].origin.y, ms[0].range, ms[0].theta, ms[200].origin.x, ms[200].origin.y,
                    ms[200].range, ms[200].theta, ms.size());

        // ActionModel::updateAction known answer (B8) and the one-particle applyAction
        ActionModel am;
        pose_xyt_t o0, o1;
        o1.x = 0.02f; o1.y = 0.01f; o1.theta = 0.01f; o1.utime = 5;
        am.updateAction(o0);
        const bool moved = am.updateAction(o1);
        const particle_t moved_p = am.applyAction(kp[1]);
        std::printf(", \"action\": [%d, %.17g, %.17g, %.17g], \"applied\": [%.9g, %.9g, %.9g, %.9g, %lld]", (int)moved,
                    am.action().rot1, am.action().trans, am.action().rot2, moved_p.pose.x, moved_p.pose.y,
                    moved_p.pose.theta, moved_p.parent_pose.x, (long long)moved_p.pose.utime);

        // the filter, driven like OccupancyGridSLAM::updateLocalization
        ParticleFilter pf(n);
        pf.setSeed(7);
        pf.setMaxExportedParticles(1000);
        pose_xyt_t pose;
        pose.x = 0.0f; pose.y = 0.0f; pose.theta = 0.0f; pose.utime = 1000000;
        pf.initializeFilterAtPose(pose);                          // slam.cpp:246
        pf.updateFilter(pose, marchScan(map, pose.x, pose.y, pose.theta, pose.utime), map);   // latches odometry
        std::printf(", \"track\": [");
        for (int k = 0; k < steps; ++k) {
            pose.x += 0.05f * std::cos(pose.theta); pose.y += 0.05f * std::sin(pose.theta); pose.theta += 0.03f;
            pose.utime += 100000;
            const lidar_t scan = marchScan(map, pose.x, pose.y, pose.theta, pose.utime);
            const pose_xyt_t est = pf.updateFilter(pose, scan, map);           // slam.cpp:259
            // the map changes between updates even when localising (slam.cpp:276): touch a few free cells
            for (int c = 0; c < 5; ++c) map.setLogOdds(100 + c + k, 100, (CellOdds)(-1 - c));
            const pose_xyt_t again = pf.poseEstimate();
            std::printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %lld, %d]", k ? ", " : "", pose.x, pose.y, pose.theta,
                        est.x, est.y, est.theta, (long long)est.utime,
                        (int)(again.x == est.x && again.y == est.y && again.theta == est.theta));
        }
        const particles_t cloud = pf.particles();                 // slam.cpp:265
        double wsum = 0;
        for (const auto& p : cloud.particles) wsum += p.weight;
        const mcl_stats st = pf.stats();
        std::printf("], \"exported\": %d, \"exported_weight\": %.9g, \"updates\": %lld, \"evals\": %lld, "
                    "\"launches\": %d, \"ms_total\": %.4f",
                    cloud.num_particles, wsum, (long long)st.updates, (long long)st.evals, st.kernel_launches,
                    st.ms_total);
        // several mirrors of one grid: the filter's engine and the stand-alone SensorModel's engine both
        // mirror `map`; writes one of them has already picked up must still reach the other, and a grid assigned from
        // another grid must be re-mirrored even though both were constructed alike
        {
            for (int y = 96; y < 104; ++y)
                for (int x = 112; x < 120; ++x) map.setLogOdds(x, y, (CellOdds)90);      // a block 0.6 m in front of (0, 0)
            pose.x += 0.05f; pose.utime += 100000;
            pf.updateFilter(pose, marchScan(map, pose.x, pose.y, pose.theta, pose.utime), map);   // the filter's mirror syncs first
            SensorModel fresh;                                                       // full upload of the current grid
            const double incremental = sm.likelihood(kp[0], ks, map), full = fresh.likelihood(kp[0], ks, map);
            OccupancyGrid a(10.0f, 10.0f, 0.05f), b(10.0f, 10.0f, 0.05f);
            for (int x = 20; x < 180; ++x) b.setLogOdds(x, 130, (CellOdds)70);
            SensorModel follower;
            const double before = follower.likelihood(kp[0], ks, a);
            a = b;                                                                   // same object, new content
            const double after = follower.likelihood(kp[0], ks, a), want = fresh.likelihood(kp[0], ks, b);
            std::printf(", \"mirrors\": [%.17g, %.17g, %.17g, %.17g, %.17g]", incremental, full, before, after, want);
        }
        // action-only mode
        pose.x += 0.05f; pose.utime += 100000;
        const pose_xyt_t ao = pf.updateFilterActionOnly(pose);
        std::printf(", \"action_only\": [%.9g, %.9g]}\n", ao.x, pose.x);
    } catch (const b200::EngineError& e) {
        std::fprintf(stderr, "engine error %d: %s\n", e.code(), e.what());
        return 10;
    }
    return 0;
}
