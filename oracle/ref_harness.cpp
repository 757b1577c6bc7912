// ORACLE / TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path (botlab_b200/).
//
// C-ABI harness around the UNMODIFIED reference sources of the Monte Carlo localization path.  The five reference
// translation units (src/slam/{particle_filter,action_model,sensor_model,moving_laser_scan,occupancy_grid}.cpp) are
// compiled where they lie under /root/reference by oracle/Makefile; nothing from them is copied here.  This file only
// (1) reaches the private stage methods through `#define private public` at include time,
// (2) neutralises the reference's latent defects FROM THE OUTSIDE:
//       - initial weights 1/N integer division (particle_filter.cpp:18)       -> weights overwritten with 1.0/N
//       - uninitialised accumulator in estimatePosteriorPose (:146)           -> zero-initialising shim pose_xyt_t
//       - ActionModel::utime_ never assigned (action_model.hpp:72)            -> set through private access
// (3) records the noise draws the reference's std::mt19937 produced so they can be injected into the engine,
// (4) exposes everything as plain C functions over POD buffers for ctypes.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <deque>
#include <limits>
#include <random>
#include <string>
#include <vector>

#define private public
#include <slam/particle_filter.hpp>
#include <slam/action_model.hpp>
#include <slam/sensor_model.hpp>
#include <slam/moving_laser_scan.hpp>
#include <slam/occupancy_grid.hpp>
#include <slam/mapping.hpp>
#include <common/pose_trace.hpp>
#include <planning/obstacle_distance_grid.hpp>
#undef private
#include <lcmtypes/lidar_t.hpp>
#include <lcmtypes/occupancy_grid_t.hpp>

extern "C" {

struct ref_pose { int64_t utime; float x, y, theta; };
struct ref_particle { ref_pose pose, parent_pose; double weight; };

}  // extern "C"

static_assert(sizeof(pose_xyt_t) == 24 && sizeof(ref_pose) == 24, "pose layout");
static_assert(sizeof(particle_t) == 56 && sizeof(ref_particle) == 56, "particle layout");

namespace {

pose_xyt_t to_pose(const ref_pose& p)
{
    pose_xyt_t q;
    q.utime = p.utime; q.x = p.x; q.y = p.y; q.theta = p.theta;
    return q;
}

ref_pose from_pose(const pose_xyt_t& p)
{
    ref_pose q;
    q.utime = p.utime; q.x = p.x; q.y = p.y; q.theta = p.theta;
    return q;
}

lidar_t make_scan(const float* ranges, const float* thetas, const int64_t* times, int n)
{
    lidar_t s;
    s.utime = n > 0 ? times[n - 1] : 0;
    s.num_ranges = n;
    s.ranges.assign(ranges, ranges + n);
    s.thetas.assign(thetas, thetas + n);
    s.times.assign(times, times + n);
    s.intensities.assign(n, 0.0f);
    return s;
}

// The three draws applyAction makes (action_model.cpp:84-86), reproduced on a COPY of the generator.
void replay_draws(std::mt19937 gen, const ActionModel& a, int n, float* out)
{
    for (int i = 0; i < n; ++i) {
        out[3 * i + 0] = std::normal_distribution<>(a.rot1_, a.rot1Std_)(gen);
        out[3 * i + 1] = std::normal_distribution<>(a.trans_, a.transStd_)(gen);
        out[3 * i + 2] = std::normal_distribution<>(a.rot2_, a.rot2Std_)(gen);
    }
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- grid
void* ref_grid_new(const int8_t* cells, int w, int h, float ox, float oy, float mpc)
{
    occupancy_grid_t m;
    m.utime = 0; m.origin_x = ox; m.origin_y = oy; m.meters_per_cell = mpc;
    m.width = w; m.height = h; m.num_cells = w * h;
    m.cells.assign(cells, cells + (size_t)w * h);
    OccupancyGrid* g = new OccupancyGrid();
    g->fromLCM(m);
    return g;
}

// Same construction order as OccupancyGridSLAM: ctor (slam.cpp:23) then loadFromFile (slam.cpp:38).
void* ref_grid_load(const char* path, float w_m, float h_m, float mpc)
{
    OccupancyGrid* g = new OccupancyGrid(w_m, h_m, mpc);
    if (!g->loadFromFile(path)) { delete g; return nullptr; }
    return g;
}

void ref_grid_info(void* gp, int* w, int* h, float* ox, float* oy, float* mpc, float* cpm)
{
    OccupancyGrid* g = (OccupancyGrid*)gp;
    *w = g->widthInCells(); *h = g->heightInCells();
    *ox = g->originInGlobalFrame().x; *oy = g->originInGlobalFrame().y;
    *mpc = g->metersPerCell(); *cpm = g->cellsPerMeter();
}

void ref_grid_cells(void* gp, int8_t* out)
{
    OccupancyGrid* g = (OccupancyGrid*)gp;
    for (int y = 0; y < g->heightInCells(); ++y)
        for (int x = 0; x < g->widthInCells(); ++x)
            out[(size_t)y * g->widthInCells() + x] = g->logOdds(x, y);
}

int ref_grid_logodds(void* gp, int x, int y) { return ((OccupancyGrid*)gp)->logOdds(x, y); }

void ref_grid_free(void* gp) { delete (OccupancyGrid*)gp; }

// ---------------------------------------------------------------- sensor model
void ref_likelihood(void* gp, const ref_particle* p, int n, const float* ranges, const float* thetas,
                    const int64_t* times, int nb, double* out)
{
    OccupancyGrid* g = (OccupancyGrid*)gp;
    lidar_t scan = make_scan(ranges, thetas, times, nb);
    SensorModel sm;
    for (int i = 0; i < n; ++i) {
        particle_t q;
        q.pose = to_pose(p[i].pose); q.parent_pose = to_pose(p[i].parent_pose); q.weight = p[i].weight;
        out[i] = sm.likelihood(q, scan, *g);
    }
}

// rays_out: 4 floats per ray (origin.x, origin.y, range, theta); returns number of rays kept.
int ref_moving_scan(const float* ranges, const float* thetas, const int64_t* times, int nb, const ref_pose* begin,
                    const ref_pose* end, float* rays_out)
{
    lidar_t scan = make_scan(ranges, thetas, times, nb);
    MovingLaserScan ms(scan, to_pose(*begin), to_pose(*end));
    int k = 0;
    for (auto& r : ms) {
        rays_out[4 * k + 0] = r.origin.x; rays_out[4 * k + 1] = r.origin.y;
        rays_out[4 * k + 2] = r.range;    rays_out[4 * k + 3] = r.theta;
        ++k;
    }
    return k;
}

// ---------------------------------------------------------------- action model
void* ref_action_new(void)
{
    ActionModel* a = new ActionModel();
    a->utime_ = 0;
    return a;
}
void ref_action_free(void* ap) { delete (ActionModel*)ap; }
void ref_action_seed(void* ap, unsigned seed) { ((ActionModel*)ap)->numberGenerator_.seed(seed); }
void ref_action_set_utime(void* ap, int64_t t) { ((ActionModel*)ap)->utime_ = t; }

// out6 = rot1, trans, rot2, rot1Std, transStd, rot2Std; returns moved.
int ref_action_update(void* ap, const ref_pose* odom, double* out6)
{
    ActionModel* a = (ActionModel*)ap;
    bool moved = a->updateAction(to_pose(*odom));
    out6[0] = a->rot1_; out6[1] = a->trans_; out6[2] = a->rot2_;
    out6[3] = a->rot1Std_; out6[4] = a->transStd_; out6[5] = a->rot2Std_;
    return moved ? 1 : 0;
}

// Applies the action to n particles in order; draws_out (3n floats, may be null) receives the draws it consumed.
void ref_action_apply(void* ap, const ref_particle* in, ref_particle* out, int n, float* draws_out)
{
    ActionModel* a = (ActionModel*)ap;
    if (draws_out) replay_draws(a->numberGenerator_, *a, n, draws_out);
    for (int i = 0; i < n; ++i) {
        particle_t q;
        q.pose = to_pose(in[i].pose); q.parent_pose = to_pose(in[i].parent_pose); q.weight = in[i].weight;
        particle_t r = a->applyAction(q);
        out[i].pose = from_pose(r.pose); out[i].parent_pose = from_pose(r.parent_pose); out[i].weight = r.weight;
    }
}

// ---------------------------------------------------------------- particle filter
void* ref_pf_new(int n)
{
    ParticleFilter* pf = new ParticleFilter(n);
    pf->actionModel_.utime_ = 0;
    return pf;
}
void ref_pf_free(void* p) { delete (ParticleFilter*)p; }

void ref_pf_set_particles(void* p, const ref_particle* in, int n)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    pf->posterior_.resize(n);
    std::memcpy(pf->posterior_.data(), in, (size_t)n * sizeof(ref_particle));
}

int ref_pf_get_particles(void* p, ref_particle* out, int max_n)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    particles_t ps = pf->particles();
    int n = ps.num_particles < max_n ? ps.num_particles : max_n;
    std::memcpy(out, ps.particles.data(), (size_t)n * sizeof(ref_particle));
    return ps.num_particles;
}

// Reference init (random_device-seeded, so non-deterministic) followed by the weight fix of Appendix C.
void ref_pf_init_at_pose(void* p, const ref_pose* pose)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    pf->initializeFilterAtPose(to_pose(*pose));
    for (auto& q : pf->posterior_) q.weight = 1.0 / pf->kNumParticles_;
}

void ref_pf_set_action_utime(void* p, int64_t t) { ((ParticleFilter*)p)->actionModel_.utime_ = t; }
void ref_pf_seed_action(void* p, unsigned seed) { ((ParticleFilter*)p)->actionModel_.numberGenerator_.seed(seed); }

// The uniform draw the reference will make after srand(seed): rand()/RAND_MAX * (1/N)  (particle_filter.cpp:89-92).
double ref_resample_draw(unsigned seed, int n)
{
    srand(seed);
    double m_inv = 1.0 / n;
    return (((double)rand()) / (double)RAND_MAX) * m_inv;
}

// Runs the reference's resamplePosteriorDistribution on the filter's current posterior_ with rand() re-seeded.
// Indices are recovered by tagging pose.utime with the source index for the duration of the call.
// Guard for Appendix C "resample can overrun": a sentinel particle with weight +inf is appended past the end so the
// reference's unbounded `while (U > c)` loop stops there; an index == n in idx_out reports the overrun to the caller.
void ref_pf_resample(void* p, unsigned seed, int32_t* idx_out)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    int n = pf->kNumParticles_;
    std::vector<int64_t> saved(n);
    for (int i = 0; i < n; ++i) { saved[i] = pf->posterior_[i].pose.utime; pf->posterior_[i].pose.utime = i; }
    particle_t sentinel;
    sentinel.pose.utime = n;
    sentinel.weight = std::numeric_limits<double>::infinity();
    pf->posterior_.push_back(sentinel);
    srand(seed);
    std::vector<particle_t> prior = pf->resamplePosteriorDistribution();
    pf->posterior_.pop_back();
    for (int m = 0; m < n; ++m) idx_out[m] = (int32_t)prior[m].pose.utime;
    for (int i = 0; i < n; ++i) pf->posterior_[i].pose.utime = saved[i];
}

// proposal -> normalised posterior (particle_filter.cpp:116-141); also leaves it in posterior_.
void ref_pf_normalize(void* p, void* gp, const ref_particle* proposal, int n, const float* ranges, const float* thetas,
                      const int64_t* times, int nb, ref_particle* out)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    std::vector<particle_t> prop(n);
    std::memcpy(prop.data(), proposal, (size_t)n * sizeof(ref_particle));
    lidar_t scan = make_scan(ranges, thetas, times, nb);
    std::vector<particle_t> post = pf->computeNormalizedPosterior(prop, scan, *(OccupancyGrid*)gp);
    std::memcpy(out, post.data(), (size_t)n * sizeof(ref_particle));
    pf->posterior_ = post;
}

void ref_pf_estimate(void* p, const ref_particle* in, int n, ref_pose* out)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    std::vector<particle_t> post(n);
    std::memcpy(post.data(), in, (size_t)n * sizeof(ref_particle));
    *out = from_pose(pf->estimatePosteriorPose(post));
}

// One full updateFilter (particle_filter.cpp:37-52) with rand() re-seeded.  draws_out (3N floats, may be null) gets the
// action-model draws in particle order; seconds_out (may be null) the steady_clock time of updateFilter alone.
// Returns whether the action model reported motion.  action_utime: value planted in ActionModel::utime_ beforehand.
int ref_pf_update(void* p, void* gp, const ref_pose* odom, const float* ranges, const float* thetas,
                  const int64_t* times, int nb, unsigned seed, int64_t action_utime, ref_pose* pose_out,
                  float* draws_out, double* seconds_out)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    lidar_t scan = make_scan(ranges, thetas, times, nb);
    pf->actionModel_.utime_ = action_utime;
    std::mt19937 gen_before = pf->actionModel_.numberGenerator_;
    {
        // overrun guard, see ref_pf_resample
        particle_t sentinel;
        sentinel.weight = std::numeric_limits<double>::infinity();
        pf->posterior_.push_back(sentinel);
    }
    srand(seed);
    auto t0 = std::chrono::steady_clock::now();
    pose_xyt_t est = pf->updateFilter(to_pose(*odom), scan, *(OccupancyGrid*)gp);
    auto t1 = std::chrono::steady_clock::now();
    bool moved = pf->actionModel_.moved_;
    if (!moved) pf->posterior_.pop_back();   // untouched posterior_ still carries the sentinel
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (draws_out && moved) replay_draws(gen_before, pf->actionModel_, pf->kNumParticles_, draws_out);
    *pose_out = from_pose(est);
    return moved ? 1 : 0;
}

int ref_pf_update_action_only(void* p, const ref_pose* odom, int64_t action_utime, ref_pose* pose_out, float* draws_out)
{
    ParticleFilter* pf = (ParticleFilter*)p;
    pf->actionModel_.utime_ = action_utime;
    std::mt19937 gen_before = pf->actionModel_.numberGenerator_;
    pose_xyt_t est = pf->updateFilterActionOnly(to_pose(*odom));
    bool moved = pf->actionModel_.moved_;
    if (draws_out && moved) replay_draws(gen_before, pf->actionModel_, pf->kNumParticles_, draws_out);
    *pose_out = from_pose(est);
    return moved ? 1 : 0;
}

void ref_pf_pose_estimate(void* p, ref_pose* out) { *out = from_pose(((ParticleFilter*)p)->poseEstimate()); }

void ref_pf_action_params(void* p, double* out6)
{
    ActionModel& a = ((ParticleFilter*)p)->actionModel_;
    out6[0] = a.rot1_; out6[1] = a.trans_; out6[2] = a.rot2_;
    out6[3] = a.rot1Std_; out6[4] = a.transStd_; out6[5] = a.rot2Std_;
}

// ---------------------------------------------------------------- headless replay of OccupancyGridSLAM's loop
// The reference's slam.cpp cannot be compiled here (it needs LCM), so its data flow (slam.cpp:88-291) is restated
// around the reference's OWN ParticleFilter, Mapping, PoseTrace, MovingLaserScan and OccupancyGrid objects.  Same
// signature as b200_replay_run (botlab_b200/src/slam/replay_c_api.cpp); noise_io receives the draws each update made.
int ref_replay_run(const int8_t* cells, int w, int h, float ox, float oy, float mpc, int have_map, int num_particles,
                   int mode, int hit_odds, int miss_odds, float max_laser_distance, int num_scans,
                   const int32_t* scan_offsets, const float* ranges, const float* thetas, const int64_t* times,
                   int num_odom, const int64_t* odom_utime, const float* odom_xyt, const float* initial_pose3,
                   unsigned rand_seed, const void* init_cloud, float* noise_io, float* poses_out,
                   int8_t* final_map_out, int* iterations_out, char* err, int err_len)
{
    (void)err; (void)err_len;
    ParticleFilter filter(num_particles);
    OccupancyGrid map(10.0f, 10.0f, 0.05f);                          // slam.cpp:23
    Mapping mapper(max_laser_distance, (int8_t)hit_odds, (int8_t)miss_odds);
    PoseTrace odometryPoses;
    std::deque<lidar_t> incoming;
    bool haveMap = false, haveInitializedPoses = false;
    if (have_map) {
        occupancy_grid_t m;
        m.utime = 0; m.origin_x = ox; m.origin_y = oy; m.meters_per_cell = mpc; m.width = w; m.height = h;
        m.num_cells = w * h;
        m.cells.assign(cells, cells + (size_t)w * h);
        map.fromLCM(m);
        haveMap = true;
    }
    pose_xyt_t initialPose, previousPose, currentPose, currentOdometry;
    initialPose.x = initial_pose3[0]; initialPose.y = initial_pose3[1]; initialPose.theta = initial_pose3[2];
    srand(rand_seed);
    int io = 0, is = 0, iter = 0;
    while (io < num_odom || is < num_scans) {
        const int64_t to = io < num_odom ? odom_utime[io] : INT64_MAX;
        const int64_t ts = is < num_scans ? times[scan_offsets[is + 1] - 1] : INT64_MAX;
        if (to <= ts) {                                              // handleOdometry, slam.cpp:131-141
            pose_xyt_t o;
            o.utime = odom_utime[io]; o.x = odom_xyt[3 * io]; o.y = odom_xyt[3 * io + 1]; o.theta = odom_xyt[3 * io + 2];
            odometryPoses.addPose(o);
            ++io;
        } else {                                                     // handleLaser, slam.cpp:90-128
            const int a = scan_offsets[is], b = scan_offsets[is + 1];
            lidar_t s = make_scan(ranges + a, thetas + a, times + a, b - a);
            if (!odometryPoses.empty() && odometryPoses.front().utime <= s.times.front()) incoming.push_back(s);
            ++is;
        }
        // isReadyToUpdate, slam.cpp:163-188
        while (!incoming.empty() && odometryPoses.containsPoseAtTime(incoming.front().times.front())) {
            lidar_t scan = incoming.front();                         // copyDataForSLAMUpdate, slam.cpp:210-229
            incoming.pop_front();
            currentOdometry = odometryPoses.poseAt(scan.times.back());
            if (!haveInitializedPoses) {                             // slam.cpp:232-250
                previousPose = initialPose;
                previousPose.utime = scan.times.front();
                currentPose = previousPose;
                currentPose.utime = scan.times.back();
                haveInitializedPoses = true;
                filter.initializeFilterAtPose(previousPose);
                if (init_cloud) std::memcpy(filter.posterior_.data(), init_cloud, sizeof(ref_particle) * num_particles);
                else for (auto& q : filter.posterior_) q.weight = 1.0 / num_particles;
            }
            bool ok = scan.num_ranges > 100;                         // slam.cpp:197
            if (ok) {
                if (haveMap) {                                       // updateLocalization, slam.cpp:253-271
                    previousPose = currentPose;
                    // the intended ActionModel::utime_ (never assigned in the reference): the odometry's utime
                    filter.actionModel_.utime_ = currentOdometry.utime;
                    std::mt19937 gen_before = filter.actionModel_.numberGenerator_;
                    particle_t sentinel;
                    sentinel.weight = std::numeric_limits<double>::infinity();
                    if (mode == 2) {
                        currentPose = filter.updateFilterActionOnly(currentOdometry);
                    } else {
                        filter.posterior_.push_back(sentinel);       // overrun guard, see ref_pf_resample
                        currentPose = filter.updateFilter(currentOdometry, scan, map);
                        if (!filter.actionModel_.moved_) filter.posterior_.pop_back();
                    }
                    if (noise_io && iter < num_scans && filter.actionModel_.moved_)
                        replay_draws(gen_before, filter.actionModel_, num_particles,
                                     noise_io + (size_t)iter * num_particles * 3);
                }
                mapper.updateMap(scan, currentPose, map);            // updateMap, slam.cpp:274-281 (always true)
                haveMap = true;
            }
            if (iter < num_scans) {
                poses_out[5 * iter + 0] = currentPose.x; poses_out[5 * iter + 1] = currentPose.y;
                poses_out[5 * iter + 2] = currentPose.theta; poses_out[5 * iter + 3] = ok ? 1.0f : 0.0f;
                poses_out[5 * iter + 4] = (float)(currentPose.utime % 1000000000LL) * 1e-6f;
            }
            ++iter;
        }
    }
    *iterations_out = iter;
    for (int y = 0; y < map.heightInCells() && y < h; ++y)
        for (int x = 0; x < map.widthInCells() && x < w; ++x) final_map_out[(size_t)y * w + x] = map.logOdds(x, y);
    return 0;
}

// ---- Mapping::updateMap (mapping.cpp:17-40) on a reference OccupancyGrid, with the private state injected: the
// previous pose and the initialized_ latch (false = the reference's first call, which changes no cell).
void ref_map_update(void* gp, const ref_pose* previous, const ref_pose* pose, int initialized, const float* ranges,
                    const float* thetas, const int64_t* times, int nb, float max_laser_distance, int hit_odds,
                    int miss_odds)
{
    Mapping mapper(max_laser_distance, (int8_t)hit_odds, (int8_t)miss_odds);
    mapper.previousPose_ = to_pose(*previous);
    mapper.initialized_ = initialized != 0;
    const lidar_t scan = make_scan(ranges, thetas, times, nb);
    mapper.updateMap(scan, to_pose(*pose), *(OccupancyGrid*)gp);
}

// ---- ObstacleDistanceGrid::setDistances (planning/obstacle_distance_grid.cpp:73-92) of a reference OccupancyGrid ----
void ref_distance_grid(void* gp, float* out)
{
    const OccupancyGrid& grid = *(OccupancyGrid*)gp;
    ObstacleDistanceGrid dist;
    dist.setDistances(grid);
    for (int y = 0; y < grid.heightInCells(); ++y)
        for (int x = 0; x < grid.widthInCells(); ++x) out[(size_t)y * grid.widthInCells() + x] = dist(x, y);
}

int ref_sizeof_particle(void) { return (int)sizeof(particle_t); }
int ref_sizeof_pose(void) { return (int)sizeof(pose_xyt_t); }
int ref_rand_max(void) { return RAND_MAX; }

}  // extern "C"
