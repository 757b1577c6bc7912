/* ORACLE / TEST INFRASTRUCTURE ONLY -- plain-C restatement of botLab's Monte Carlo localization update.
 * See mcl_oracle.h.  Every function names the reference lines it restates (paths relative to the reference root).
 * Build: gcc -O2 -ffp-contract=off (no -march, no fast-math) so the arithmetic is SSE2 scalar without FMA contraction,
 * like the reference build (src/common.mk:24-29, src/slam/Makefile:4-12).  Mixed float/double evaluation order below is
 * part of the contract: do not "simplify". */
#define _GNU_SOURCE
#include "mcl_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- common/angle_functions.hpp ------------------------------------------------------------------------------ */

/* angle_functions.hpp:12-24: float accumulator, comparisons and the 2*pi step in double. */
float orc_wrap_to_pi(float a)
{
    if ((double)a < -M_PI) {
        while ((double)a < -M_PI) a = (float)((double)a + 2.0 * M_PI);
    } else if ((double)a > M_PI) {
        while ((double)a > M_PI) a = (float)((double)a - 2.0 * M_PI);
    }
    return a;
}

/* angle_functions.hpp:78-87 */
double orc_angle_diff(double l, double r)
{
    double d = l - r;
    if (fabs(d) > M_PI) d -= (d > 0) ? M_PI * 2 : M_PI * -2;
    return d;
}

/* angle_functions.hpp:128-138 */
double orc_angle_sum(double a, double b)
{
    double s = a + b;
    if (fabs(s) > M_PI) s -= (s > 0) ? M_PI * 2 : M_PI * -2;
    return s;
}

/* ---- common/interpolation.hpp:24-50 ------------------------------------------------------------------------- */
void orc_interpolate_pose(int64_t t, const orc_pose* before, const orc_pose* after, orc_pose* out)
{
    if (before->utime == after->utime) {           /* :29-34 equal-time early-out returns `after` */
        *out = *after;
        out->utime = t;
        return;
    }
    double ratio = (double)(t - before->utime) / (double)(after->utime - before->utime);   /* :36 */
    double xs = (double)(after->x - before->x) * ratio;          /* float subtraction, double product  :39 */
    double ys = (double)(after->y - before->y) * ratio;          /* :40 */
    double ts = orc_angle_diff((double)after->theta, (double)before->theta) * ratio;        /* :41 */
    out->utime = t;
    out->x = (float)((double)before->x + xs);                    /* :45 */
    out->y = (float)((double)before->y + ys);                    /* :46 */
    out->theta = (float)orc_angle_sum((double)before->theta, ts);/* :47 */
}

/* ---- slam/moving_laser_scan.cpp:8-39 --------------------------------------------------------------------------
 * rays_out4: origin.x, origin.y, range, theta per kept ray.  Returns the number kept. */
int orc_moving_scan(const float* ranges, const float* thetas, const int64_t* times, int nb, const orc_pose* begin,
                    const orc_pose* end, float* rays_out4)
{
    int k = 0;
    for (int n = 0; n < nb; ++n) {
        if (ranges[n] > 0.15f) {                                  /* :24 */
            orc_pose rp;
            orc_interpolate_pose(times[n], begin, end, &rp);      /* :26 */
            rays_out4[4 * k + 0] = rp.x;
            rays_out4[4 * k + 1] = rp.y;
            rays_out4[4 * k + 2] = ranges[n];
            rays_out4[4 * k + 3] = orc_wrap_to_pi(rp.theta - thetas[n]);   /* float subtraction  :33 */
            ++k;
        }
    }
    return k;
}

/* ---- slam/occupancy_grid.cpp:55-71 ----------------------------------------------------------------------------- */
int orc_logodds(const orc_grid* g, int x, int y)
{
    if (x >= 0 && x < g->width && y >= 0 && y < g->height) return g->cells[(size_t)y * g->width + x];
    return 0;
}

/* float -> int as the x86-64 host does it (cvttss2si): truncation; out-of-range and NaN give INT_MIN. */
static int f2i(float v)
{
    if (!(v > -2147483648.0f && v < 2147483648.0f)) return INT_MIN;
    return (int)v;
}

/* slam/sensor_model.cpp:61-86 -- exactly one Bresenham step from (x1,y1) toward (x2,y2), then read that cell.
 * Differences are formed in wrapping unsigned arithmetic so INT_MIN inputs are defined (the reference overflows). */
static int step_odds(const orc_grid* g, int x1, int y1, int x2, int y2)
{
    int dx = (int)((unsigned)x2 - (unsigned)x1);
    int dy = (int)((unsigned)y2 - (unsigned)y1);
    if (dx < 0) dx = (int)(0u - (unsigned)dx);
    if (dy < 0) dy = (int)(0u - (unsigned)dy);
    int sx = x1 < x2 ? 1 : -1;
    int sy = y1 < y2 ? 1 : -1;
    int err = (int)((unsigned)dx - (unsigned)dy);
    double e2 = 2.0 * (double)err;
    int x = x1, y = y1;
    if (e2 >= -(double)dy) x = (int)((unsigned)x + (unsigned)sx);
    if (e2 <= (double)dx) y = (int)((unsigned)y + (unsigned)sy);
    return orc_logodds(g, x, y);
}

/* slam/sensor_model.cpp:28-59 with common/grid_utils.hpp:50-55.  *gathers += number of map reads made. */
double orc_score_ray(const orc_grid* g, float ox, float oy, float range, float theta, int* gathers)
{
    /* rayStart: double math, stored in Point<float> */
    float sx = (float)(((double)ox - (double)g->origin_x) * (double)g->cells_per_meter);
    float sy = (float)(((double)oy - (double)g->origin_y) * (double)g->cells_per_meter);
    float s, c;
    sincosf(theta, &s, &c);          /* g++ -O3 merges std::cos/std::sin(float) into one sincosf call */
    float cpm = g->cells_per_meter;
    int ex = f2i((range * c) * cpm + sx);                         /* :34 */
    int ey = f2i((range * s) * cpm + sy);                         /* :35 */
    int xx = f2i(((2 * range) * c) * cpm + sx);                   /* :37 */
    int xy = f2i(((2 * range) * s) * cpm + sy);                   /* :38 */
    double odds = orc_logodds(g, ex, ey);                         /* :41 */
    if (gathers) *gathers += 1;
    if (odds > 0) return odds;
    odds = 0;
    int o1 = step_odds(g, ex, ey, f2i(sx), f2i(sy));              /* :48 toward the robot */
    int o2 = step_odds(g, ex, ey, xx, xy);                        /* :49 away from the robot */
    if (gathers) *gathers += 2;
    if (o1 > 0) odds += 0.5 * o1;
    else if (o2 > 0) odds += 0.5 * o2;
    return odds;
}

/* slam/sensor_model.cpp:14-25 for n particles.  gathers_out/evals_out (nullable) accumulate map reads and rays. */
void orc_likelihood(const orc_grid* g, const orc_particle* p, int n, const float* ranges, const float* thetas,
                    const int64_t* times, int nb, double* out, int64_t* gathers_out, int64_t* evals_out)
{
    float* rays = (float*)malloc(sizeof(float) * 4 * (size_t)(nb > 0 ? nb : 1));
    int64_t G = 0, E = 0;
    for (int i = 0; i < n; ++i) {
        int k = orc_moving_scan(ranges, thetas, times, nb, &p[i].parent_pose, &p[i].pose, rays);
        double score = 0.0;
        int gth = 0;
        for (int j = 0; j < k; ++j)
            score += orc_score_ray(g, rays[4 * j], rays[4 * j + 1], rays[4 * j + 2], rays[4 * j + 3], &gth);
        out[i] = score;
        G += gth;
        E += k;
    }
    if (gathers_out) *gathers_out += G;
    if (evals_out) *evals_out += E;
    free(rays);
}

/* ---- slam/action_model.cpp ------------------------------------------------------------------------------------ */
void orc_action_init(orc_action* a)
{
    memset(a, 0, sizeof(*a));
}

/* action_model.cpp:22-75 */
int orc_action_update(orc_action* a, const orc_pose* odom)
{
    if (!a->initialized) { a->prev = *odom; a->initialized = 1; }
    float dx = odom->x - a->prev.x;                                /* :29 float */
    float dy = odom->y - a->prev.y;
    float dth = (float)orc_angle_diff((double)odom->theta, (double)a->prev.theta);   /* :31 */
    float dir = 1.0f;
    a->rot1 = orc_angle_diff((double)atan2f(dy, dx), (double)a->prev.theta);          /* :34 std::atan2(float,float) */
    a->trans = (double)sqrtf(dx * dx + dy * dy);                   /* :35 std::sqrt(float) */
    if (fabs(a->trans) < 0.0001) {
        a->rot1 = 0.0;
    } else if (fabs(a->rot1) > M_PI / 2.0) {                       /* :40-43 backward motion */
        a->rot1 = -orc_angle_diff(M_PI, a->rot1);
        dir = -1.0f;
    }                                                              /* :44-47 is unreachable */
    a->trans *= (double)dir;
    a->rot2 = orc_angle_diff((double)dth, a->rot1);                /* :50 */
    a->moved = !((fabs(a->trans) + fabs(a->rot2)) < (double)0.00001f);   /* :52 */
    a->rot1_std = 0.05; a->trans_std = 0.005; a->rot2_std = 0.05;  /* :64-66 */
    a->prev = *odom;
    return a->moved;
}

/* action_model.cpp:78-103 with the three float draws (rot1, trans, rot2 per particle) injected. */
void orc_action_apply(const orc_action* a, int64_t utime, const orc_particle* in, orc_particle* out, int n,
                      const float* draws3n)
{
    for (int i = 0; i < n; ++i) {
        orc_particle q = in[i];
        if (a->moved) {
            float r1 = draws3n[3 * i + 0], tr = draws3n[3 * i + 1], r2 = draws3n[3 * i + 2];
            float th = in[i].pose.theta;
            float h = th + r1;                                                  /* float sum */
            q.pose.x = (float)((double)q.pose.x + (double)tr * cos((double)h)); /* :88 unqualified cos -> double */
            q.pose.y = (float)((double)q.pose.y + (double)tr * sin((double)h)); /* :89 */
            q.pose.theta = orc_wrap_to_pi((th + r1) + r2);                      /* :90 float sums, left to right */
        }
        q.pose.utime = utime;                                                   /* :92/:99 */
        q.parent_pose = in[i].pose;                                             /* :93/:100 */
        out[i] = q;
    }
}

/* ---- slam/particle_filter.cpp --------------------------------------------------------------------------------- */

/* particle_filter.cpp:120-138: floor at 0.001, sequential double sum, divide. */
void orc_normalize(const double* scores, int n, double* weights_out, double* wsum_out)
{
    double wsum = 0.0;
    for (int i = 0; i < n; ++i) {
        double w = scores[i];
        if (w < 0.001) w = 0.001;
        weights_out[i] = w;
        wsum += w;
    }
    for (int i = 0; i < n; ++i) weights_out[i] /= wsum;
    if (wsum_out) *wsum_out = wsum;
}

/* particle_filter.cpp:84-103 with r (= rand()/RAND_MAX/N) injected.  The reference's while loop has no bound on i;
 * here i stops at n-1 and the return value counts draws that would have run past the end. */
int orc_resample(const double* weights, int n, double r, int32_t* idx_out)
{
    int i = 0, overruns = 0;
    double m_inv = 1.0 / n;
    double c = weights[0];
    for (int m = 0; m < n; ++m) {
        double u = r + m * m_inv;
        while (u > c) {
            if (i == n - 1) { ++overruns; break; }
            ++i;
            c += weights[i];
        }
        idx_out[m] = i;
    }
    return overruns;
}

/* particle_filter.cpp:144-160 (accumulator zero-initialised): x,y are FLOAT running sums, sin/cos are float,
 * their weighted sums double, atan2 double. */
void orc_estimate(const orc_particle* p, int n, orc_pose* out)
{
    float x = 0.0f, y = 0.0f;
    double ws = 0.0, wc = 0.0;
    for (int i = 0; i < n; ++i) {
        x = (float)((double)x + p[i].weight * (double)p[i].pose.x);
        y = (float)((double)y + p[i].weight * (double)p[i].pose.y);
        float s, c;
        sincosf(p[i].pose.theta, &s, &c);
        ws += p[i].weight * (double)s;
        wc += p[i].weight * (double)c;
    }
    out->x = x; out->y = y;
    out->theta = (float)atan2(ws, wc);
}

/* ---- libstdc++ 13 <random> as used by the reference (third-party, not vendored) ------------------------------
 * std::mt19937 (Matsumoto & Nishimura MT19937, 32-bit), generate_canonical<double,53> (bits/random.tcc:3349-3381) and
 * std::normal_distribution<double> (Marsaglia polar, bits/random.tcc:1811-1844). */
void orc_rng_seed(orc_rng* g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
    g->saved_available = 0;
    g->saved = 0.0;
}

uint32_t orc_rng_next(orc_rng* g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static double canonical53(orc_rng* g)
{
    double lo = (double)orc_rng_next(g);
    double hi = (double)orc_rng_next(g);
    double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* fresh_distribution != 0: a new distribution object per draw, as applyAction constructs one per call
 * (action_model.cpp:84-86), so the cached second variate is never used.  == 0: persistent object
 * (particle_filter.cpp:23-28). */
double orc_rng_normal(orc_rng* g, double mean, double stddev, int fresh_distribution)
{
    double ret;
    if (fresh_distribution) g->saved_available = 0;
    if (g->saved_available) {
        g->saved_available = 0;
        ret = g->saved;
    } else {
        double x, y, r2;
        do {
            x = 2.0 * canonical53(g) - 1.0;
            y = 2.0 * canonical53(g) - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        double mult = sqrt(-2 * log(r2) / r2);
        g->saved = x * mult;
        g->saved_available = 1;
        ret = y * mult;
    }
    return ret * stddev + mean;
}

/* The draws applyAction would consume for n particles (action_model.cpp:84-86). */
void orc_action_draws(orc_rng* g, const orc_action* a, int n, float* draws3n)
{
    for (int i = 0; i < n; ++i) {
        draws3n[3 * i + 0] = (float)orc_rng_normal(g, a->rot1, a->rot1_std, 1);
        draws3n[3 * i + 1] = (float)orc_rng_normal(g, a->trans, a->trans_std, 1);
        draws3n[3 * i + 2] = (float)orc_rng_normal(g, a->rot2, a->rot2_std, 1);
    }
}

/* particle_filter.cpp:16-34 with a given generator and the intended weight 1.0/N. */
void orc_init_at_pose(orc_rng* g, const orc_pose* pose, orc_particle* out, int n)
{
    g->saved_available = 0;
    for (int i = 0; i < n; ++i) {
        out[i].pose.x = (float)((double)pose->x + orc_rng_normal(g, 0.0, 0.01, 0));
        out[i].pose.y = (float)((double)pose->y + orc_rng_normal(g, 0.0, 0.01, 0));
        out[i].pose.theta = orc_wrap_to_pi((float)((double)pose->theta + orc_rng_normal(g, 0.0, 0.01, 0)));
        out[i].pose.utime = pose->utime;
        out[i].parent_pose = out[i].pose;
        out[i].weight = 1.0 / n;
    }
    out[n - 1].pose = *pose;                                     /* :33 (parent_pose keeps the sampled value) */
}

/* particle_filter.cpp:37-52 */
int orc_update(orc_action* a, const orc_grid* g, orc_particle* particles, orc_particle* scratch, int n,
               const orc_pose* odom, int64_t action_utime, const float* ranges, const float* thetas,
               const int64_t* times, int nb, double r, const float* draws3n, orc_pose* pose_io)
{
    int moved = orc_action_update(a, odom);
    if (moved) {
        double* w = (double*)malloc(sizeof(double) * (size_t)n);
        int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
        for (int i = 0; i < n; ++i) w[i] = particles[i].weight;
        orc_resample(w, n, r, idx);
        for (int m = 0; m < n; ++m) scratch[m] = particles[idx[m]];           /* prior */
        orc_action_apply(a, action_utime, scratch, particles, n, draws3n);    /* proposal */
        orc_likelihood(g, particles, n, ranges, thetas, times, nb, w, 0, 0);
        orc_normalize(w, n, w, 0);
        for (int i = 0; i < n; ++i) particles[i].weight = w[i];
        orc_estimate(particles, n, pose_io);
        free(w);
        free(idx);
    }
    pose_io->utime = odom->utime;
    return moved;
}

/* ---- slam/mapping.cpp:17-127 -- Mapping::updateMap on a writable grid ------------------------------------------
 * cells is modified in place.  previous / initialized are Mapping's private state (previousPose_, initialized_):
 * until the first call has latched a previous pose no cell changes (mapping.cpp:73-76, 87-90).  Occupied endpoints
 * are raised first (:26-30), then every ray lowers the cells of its Bresenham walk from the start cell up to, not
 * including, the endpoint cell (:32-36, :101-127).  Returns the number of cell writes (diagnostic). */
static void orc_map_raise(int8_t* c, int hit)            /* mapping.cpp:73-85 */
{
    if (127 - (int)*c > hit) *c = (int8_t)((int)*c + hit);
    else *c = 127;
}
static void orc_map_lower(int8_t* c, int miss)           /* mapping.cpp:87-99 */
{
    if ((int)*c - miss > -128) *c = (int8_t)((int)*c - miss);
    else *c = -128;
}
long orc_map_update(int8_t* cells, int32_t width, int32_t height, float origin_x, float origin_y, float cells_per_meter,
                    const orc_pose* previous, const orc_pose* pose, int initialized, const float* ranges,
                    const float* thetas, const int64_t* times, int nb, float max_laser_distance, int hit_odds,
                    int miss_odds)
{
    if (nb <= 0) return 0;
    orc_pose prev = initialized ? *previous : *pose;                      /* :19-21 */
    float* rays = (float*)malloc(sizeof(float) * 4 * (size_t)nb);
    const int nr = orc_moving_scan(ranges, thetas, times, nb, &prev, pose, rays);   /* :22 */
    long writes = 0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = 0; k < nr; ++k) {
            const float ox = rays[4 * k], oy = rays[4 * k + 1], range = rays[4 * k + 2], theta = rays[4 * k + 3];
            if (!(range <= max_laser_distance)) continue;                 /* :43, :60 */
            const float sx = (float)(((double)ox - (double)origin_x) * (double)cells_per_meter);   /* grid_utils.hpp:50-55 */
            const float sy = (float)(((double)oy - (double)origin_y) * (double)cells_per_meter);
            float s, c;
            sincosf(theta, &s, &c);
            const int cx = f2i((range * c) * cells_per_meter + sx);       /* :48, :65 */
            const int cy = f2i((range * s) * cells_per_meter + sy);
            if (pass == 0) {                                              /* scoreEndpoint :42-57 */
                if (initialized && cx >= 0 && cx < width && cy >= 0 && cy < height) {
                    orc_map_raise(&cells[(size_t)cy * width + cx], hit_odds);
                    ++writes;
                }
            } else {                                                      /* scoreRay -> bresenham :101-127 */
                int x = f2i(sx), y = f2i(sy);
                const int dx = abs(cx - x), dy = abs(cy - y);
                const int stepx = x < cx ? 1 : -1, stepy = y < cy ? 1 : -1;
                int err = dx - dy;
                while (x != cx || y != cy) {
                    if (initialized && x >= 0 && x < width && y >= 0 && y < height) {
                        orc_map_lower(&cells[(size_t)y * width + x], miss_odds);
                        ++writes;
                    }
                    const float e2 = (float)(2 * err);
                    if (e2 >= (float)-dy) { err -= dy; x += stepx; }
                    if (e2 <= (float)dx) { err += dx; y += stepy; }
                }
            }
        }
    }
    free(rays);
    return writes;
}

/* Per-ray scores of ONE particle, in scan order of its valid beams (what orc_likelihood sums): test helper for checks
 * that need to know which evaluation contributed what.  Returns the number of valid beams written to out. */
int orc_ray_scores(const orc_grid* g, const orc_particle* p, const float* ranges, const float* thetas,
                   const int64_t* times, int nb, double* out)
{
    float* rays = (float*)malloc(sizeof(float) * 4 * (size_t)(nb > 0 ? nb : 1));
    const int nr = orc_moving_scan(ranges, thetas, times, nb, &p->parent_pose, &p->pose, rays);   /* sensor_model.cpp:16 */
    for (int k = 0; k < nr; ++k)
        out[k] = orc_score_ray(g, rays[4 * k], rays[4 * k + 1], rays[4 * k + 2], rays[4 * k + 3], 0);
    free(rays);
    return nr;
}

/* ---- planning/obstacle_distance_grid.cpp:44-188 -- ObstacleDistanceGrid::setDistances ----------------------------------
 * initializeDistances: free cells (log-odds < 0) start at -1, every other cell (log-odds >= 0: occupied OR unknown) at 0
 * (:52-67).  enqueue_obstacle_cells expands every non-free cell (:133-152); expand_node visits the four neighbours and
 * gives an unvisited (-1) one the distance node.distance + 0.1f (:154-188; "should be 0.05" says the reference -- kept).
 * The priority queue orders nodes by distance and every step costs the same, so a cell is reached first along a
 * shortest four-connected path: the literal queue below (FIFO = the same order up to ties between equal distances,
 * which cannot change the value a cell gets) restates it.  thr generalises the source rule: sources are cells with
 * log-odds >= thr (thr = 0: the reference).  Returns the number of cells the brushfire never reached (left at -1). */
long orc_distance_grid(const int8_t* cells, int32_t width, int32_t height, int thr, float* out)
{
    const long n = (long)width * height;
    int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    long head = 0, tail = 0;
    for (long i = 0; i < n; ++i) out[i] = cells[i] >= thr ? 0.0f : -1.0f;
    static const int dx[4] = {1, -1, 0, 0}, dy[4] = {0, 0, 1, -1};
    for (int pass = 0; pass < 2; ++pass) {
        /* pass 0: expand every source cell in scan order (enqueue_obstacle_cells); pass 1: drain the queue */
        long count = pass == 0 ? n : 0;
        for (long i = 0; pass == 0 ? i < count : head < tail; ++i) {
            long c;
            if (pass == 0) { c = i; if (!(cells[c] >= thr)) continue; }
            else c = queue[head++];
            const int x = (int)(c % width), y = (int)(c / width);
            for (int k = 0; k < 4; ++k) {
                const int ax = x + dx[k], ay = y + dy[k];
                if (ax < 0 || ax >= width || ay < 0 || ay >= height) continue;
                const long a = (long)ay * width + ax;
                if (out[a] == -1.0f) {
                    out[a] = out[c] + 0.1f;
                    queue[tail++] = (int32_t)a;
                }
            }
        }
    }
    long unreached = 0;
    for (long i = 0; i < n; ++i) unreached += out[i] == -1.0f;
    free(queue);
    return unreached;
}

/* ---- likelihood-field sensor mode (an EXTENSION of the engine, mcl_params.sensor_mode = 1; not in the reference) ---------
 * field u(cell) = max(0, 127 - 8 d^2), d = four-connected steps to the nearest OCCUPIED cell (log-odds > 0), i.e. the
 * brushfire above with thr = 1; a ray scores u at the reference's own endpoint cell (sensor_model.cpp:34-35: the
 * truncated float coordinates), 0 outside the grid; a particle's score is the sum over its valid rays. */
static int lf_value_from_distance(float d)
{
    if (d < 0.0f) return 0;
    const int steps = (int)(d * 10.0f + 0.5f);
    return steps >= 4 ? 0 : 127 - 8 * steps * steps;
}

void orc_likelihood_field(const orc_grid* g, const orc_particle* p, int n, const float* ranges, const float* thetas,
                          const int64_t* times, int nb, double* out)
{
    const long cellsn = (long)g->width * g->height;
    float* dist = (float*)malloc(sizeof(float) * (size_t)(cellsn > 0 ? cellsn : 1));
    orc_distance_grid(g->cells, g->width, g->height, 1, dist);
    float* rays = (float*)malloc(sizeof(float) * 4 * (size_t)(nb > 0 ? nb : 1));
    for (int i = 0; i < n; ++i) {
        const int k = orc_moving_scan(ranges, thetas, times, nb, &p[i].parent_pose, &p[i].pose, rays);
        double score = 0.0;
        for (int j = 0; j < k; ++j) {
            const float ox = rays[4 * j], oy = rays[4 * j + 1], range = rays[4 * j + 2], theta = rays[4 * j + 3];
            const float sx = (float)(((double)ox - (double)g->origin_x) * (double)g->cells_per_meter);
            const float sy = (float)(((double)oy - (double)g->origin_y) * (double)g->cells_per_meter);
            float sn, cs;
            sincosf(theta, &sn, &cs);
            const float cpm = g->cells_per_meter;
            const int ex = f2i((range * cs) * cpm + sx), ey = f2i((range * sn) * cpm + sy);
            if (ex >= 0 && ex < g->width && ey >= 0 && ey < g->height)
                score += (double)lf_value_from_distance(dist[(size_t)ey * g->width + ex]);
        }
        out[i] = score;
    }
    free(rays);
    free(dist);
}
