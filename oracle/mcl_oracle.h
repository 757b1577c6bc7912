/* ORACLE / TEST INFRASTRUCTURE ONLY -- plain-C restatement of botLab's Monte Carlo localization update.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this; the product (botlab_b200/)
 * never does.  Parity status: PINNED against the compiled, unmodified reference (oracle/_ref, see tests/test_oracle.py
 * and tests/golden/); the reference itself ships no tests or golden vectors for this path. */
#ifndef MCL_ORACLE_H
#define MCL_ORACLE_H
#include <stdint.h>

typedef struct { int64_t utime; float x, y, theta; } orc_pose;               /* lcmtypes/pose_xyt_t.lcm:1-8 */
typedef struct { orc_pose pose, parent_pose; double weight; } orc_particle;  /* lcmtypes/particle_t.lcm:4-9 */
typedef struct {
    const int8_t* cells;    /* row-major, index y*width+x  (occupancy_grid.hpp:208) */
    int32_t width, height;
    float origin_x, origin_y;
    float cells_per_meter;
} orc_grid;

typedef struct {            /* ActionModel state (action_model.hpp:66-76) */
    orc_pose prev;
    int initialized, moved;
    double rot1, trans, rot2, rot1_std, trans_std, rot2_std;
} orc_action;

typedef struct {            /* std::mt19937 + one std::normal_distribution<double>'s cached variate */
    uint32_t mt[624];
    int idx;
    int saved_available;
    double saved;
} orc_rng;

float  orc_wrap_to_pi(float a);
double orc_angle_diff(double l, double r);
double orc_angle_sum(double a, double b);
void   orc_interpolate_pose(int64_t t, const orc_pose* before, const orc_pose* after, orc_pose* out);
int    orc_moving_scan(const float* ranges, const float* thetas, const int64_t* times, int nb, const orc_pose* begin,
                       const orc_pose* end, float* rays_out4);
int    orc_logodds(const orc_grid* g, int x, int y);
double orc_score_ray(const orc_grid* g, float ox, float oy, float range, float theta, int* gathers);
void   orc_likelihood(const orc_grid* g, const orc_particle* p, int n, const float* ranges, const float* thetas,
                      const int64_t* times, int nb, double* out, int64_t* gathers_out, int64_t* evals_out);

void   orc_action_init(orc_action* a);
int    orc_action_update(orc_action* a, const orc_pose* odom);
void   orc_action_apply(const orc_action* a, int64_t utime, const orc_particle* in, orc_particle* out, int n,
                        const float* draws3n);

void   orc_normalize(const double* scores, int n, double* weights_out, double* wsum_out);
int    orc_resample(const double* weights, int n, double r, int32_t* idx_out);
void   orc_estimate(const orc_particle* p, int n, orc_pose* out);

void   orc_rng_seed(orc_rng* g, uint32_t seed);
uint32_t orc_rng_next(orc_rng* g);
double orc_rng_normal(orc_rng* g, double mean, double stddev, int fresh_distribution);
void   orc_action_draws(orc_rng* g, const orc_action* a, int n, float* draws3n);
void   orc_init_at_pose(orc_rng* g, const orc_pose* pose, orc_particle* out, int n);

/* One ParticleFilter::updateFilter with the resample draw r and the action draws injected.
 * particles: in = posterior of the previous update, out = new posterior.  Returns moved. */
int    orc_update(orc_action* a, const orc_grid* g, orc_particle* particles, orc_particle* scratch, int n,
                  const orc_pose* odom, int64_t action_utime, const float* ranges, const float* thetas,
                  const int64_t* times, int nb, double r, const float* draws3n, orc_pose* pose_io);
int    orc_ray_scores(const orc_grid* g, const orc_particle* p, const float* ranges, const float* thetas,
                      const int64_t* times, int nb, double* out);
/* Mapping::updateMap (mapping.cpp:17-127) on a writable int8 grid; see mcl_oracle.c. */
long   orc_map_update(int8_t* cells, int32_t width, int32_t height, float origin_x, float origin_y, float cells_per_meter,
                      const orc_pose* previous, const orc_pose* pose, int initialized, const float* ranges,
                      const float* thetas, const int64_t* times, int nb, float max_laser_distance, int hit_odds,
                      int miss_odds);
/* ObstacleDistanceGrid::setDistances (planning/obstacle_distance_grid.cpp:44-188); thr = 0 is the reference's rule. */
long   orc_distance_grid(const int8_t* cells, int32_t width, int32_t height, int thr, float* out);
/* The engine's likelihood-field sensor mode (extension): see mcl_oracle.c. */
void   orc_likelihood_field(const orc_grid* g, const orc_particle* p, int n, const float* ranges, const float* thetas,
                            const int64_t* times, int nb, double* out);
#endif
