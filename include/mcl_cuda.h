/* mcl_cuda.h -- C ABI of the B200-native Monte Carlo localization engine (libmcl_cuda.so).
 *
 * This is the thin extern-"C" layer of the new src/slam/cuda/ module: the entry points botLab's C++ host classes
 * (ParticleFilter / ActionModel / SensorModel / MovingLaserScan / OccupancyGrid device mirror) bind instead of running
 * their serial CPU loops.  The reference has no FFI of its own; each function below names the reference code it replaces
 * (paths relative to the reference root).  Plain pointers and sizes only; every pointer argument is caller-owned HOST
 * memory valid for the duration of the call; device memory is owned by the engine.
 *
 * Conventions: every function returns 0 on success or a negative MCL_ERR_* code; mcl_last_error() gives the message.
 * Nothing throws or aborts.  One caller thread per handle; one CUDA stream per handle.  There is no CPU fallback:
 * without a CUDA device mcl_create fails.
 */
#ifndef MCL_CUDA_H
#define MCL_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCL_OK 0
#define MCL_ERR_INVALID (-1)    /* bad argument */
#define MCL_ERR_CUDA (-2)       /* CUDA runtime / driver failure */
#define MCL_ERR_STATE (-3)      /* call order (no map, no particles, no scan ...) */
#define MCL_ERR_NO_DEVICE (-4)  /* no usable sm_100 GPU */
#define MCL_ERR_COMM (-5)       /* multi-GPU exchange failure */

typedef struct mcl_engine mcl_engine;

/* lcmtypes/pose_xyt_t.lcm:1-8 -- 24 bytes */
typedef struct { int64_t utime; float x, y, theta; } mcl_pose_t;
/* lcmtypes/particle_t.lcm:4-9 -- 56 bytes, identical to the reference's particle_t */
typedef struct { mcl_pose_t pose, parent_pose; double weight; } mcl_particle_t;

/* Compile-time constants of the reference exposed as parameters; defaults are the reference's values. */
typedef struct {
    float  min_range;           /* moving_laser_scan.cpp:24   0.15f  */
    double weight_floor;        /* particle_filter.cpp:121    0.001  */
    double init_std;            /* particle_filter.cpp:23     0.01   */
    int    legacy_equal_utime;  /* 1: every particle keeps pose.utime == parent_pose.utime, the reference's de-facto
                                   behaviour (ActionModel::utime_ is never assigned, action_model.hpp:72), so the
                                   per-ray interpolation degenerates to "all rays from the current pose".  0 (default):
                                   pose.utime = odometry utime, real per-ray interpolation (the evident intent). */
    int    lanes_per_particle;  /* sensor kernel mapping: 0 = auto, else 1, 2, 4, 8, 16 or 32 lanes share one particle */
    int    map_tile;            /* 0 = auto, 1 = force L2/global gathers, 2 = force shared-memory map tile */
    int    sensor_path;         /* 0 = auto: the score-table pass (one kernel: certified table lookups, the literal
                                   restatement for whatever it cannot certify) where the cloud's map window fits shared
                                   memory, else the two-pass path; 1 = literal restatement only; 2 = two-pass path only
                                   (certified float pass, then the literal restatement in a second kernel).  Results are
                                   identical whichever runs */
    int    weight_mode;         /* 0 (default) = the reference's linear rule w = max(score, weight_floor) / sum
                                   (particle_filter.cpp:116-141): the parity mode.  1 = scores as log-likelihoods:
                                   w = exp(lse_beta (score - max score)) / sum, normalised by a max / log-sum-exp reduction
                                   (an extension: it changes results, so it is never the default) */
    int    sensor_mode;         /* 0 (default) = the reference's beam-end scoring (sensor_model.cpp:28-59): the parity mode.
                                   1 = likelihood field (an extension, never the default): a ray scores u(endpoint cell),
                                   u = max(0, 127 - 8 d^2), d = four-connected steps to the nearest occupied cell -- a
                                   distance grid computed like planning/obstacle_distance_grid.cpp:73-188 (brushfire),
                                   one table lookup per beam; pairs naturally with weight_mode 1 */
    double lse_beta;            /* inverse temperature of weight_mode 1 (default 0.05 per score unit) */
    int    reserved[4];
} mcl_params;

/* ActionModel state + per-update parameters (action_model.hpp:66-76). */
typedef struct {
    mcl_pose_t previous_odometry;
    int    initialized;
    int    moved;
    double rot1, trans, rot2;
    double rot1_std, trans_std, rot2_std;
} mcl_action_t;

typedef struct {
    int64_t num_particles;        /* global particle count */
    int64_t local_particles;      /* particles scored by this rank */
    int64_t updates;              /* fused updates run so far */
    int64_t valid_beams;          /* beams of the current scan with range > min_range */
    int64_t evals;                /* particle-beam evaluations of the last mcl_score / mcl_update (this rank) */
    int64_t gathers;              /* map reads of the last scoring pass if gather counting is on, else -1 */
    int64_t resample_overruns;    /* draws the reference's unbounded loop would have run past the end for (clamped) */
    int64_t seq_fallback_chunks;  /* chunks of the exact sequential-sum emulation that took the serial path */
    double  weight_sum;           /* wSum of the last normalise (sequential-double semantics) */
    double  effective_sample_size;
    float   ms_resample, ms_action, ms_score, ms_normalize, ms_estimate, ms_total;  /* last mcl_update, CUDA events */
    int     lanes_per_particle;   /* mapping actually used by the last scoring pass */
    int     map_tile_used;        /* 1 = L2/global gathers, 2 = one shared-memory tile, 3 = one tile per batch of 1024 particles,
                                     4 = one shared-memory class tile + score table (score-table pass), 5 = the same
                                     rebuilt per batch of 4096 (or 1024) particles (dense global-localisation clouds) */
    int     kernel_launches;      /* kernels launched by the last mcl_update */
    int     collectives;          /* slice exchanges enqueued by the last mcl_update (0 on one GPU) */
    int     peer_push;            /* 1: pose slices travel by copy-engine peer writes (CUDA IPC), else NCCL all-gather */
    int     sensor_path;          /* path of the last scoring pass: 3 = score-table pass, 2 = certified float pass + exact
                                     re-evaluation (two kernels), 1 = exact only */
    int     table_variant;        /* score-table pass: 0 / 1 = one window, 16- / 8-bit classes; 2 / 3 = one window per batch */
    int     culled_beams;         /* score-table pass: beams of the last scan that score 0 for every particle of the slice
                                     (their endpoints can only lie in open space) and were not evaluated */
    int64_t deferred_evals;       /* evaluations of the last scoring pass the float pass could not certify (re-done exactly) */
    double  fast_eps;             /* error bound (cells) the certification used, 0 when the exact path ran alone */
} mcl_stats;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */
void mcl_default_params(mcl_params* p);
/* ParticleFilter::ParticleFilter(int) (particle_filter.cpp:8-13).  device: CUDA ordinal. */
int  mcl_create(const mcl_params* params_or_null, int64_t num_particles, int device, mcl_engine** out);
void mcl_destroy(mcl_engine* h);
const char* mcl_last_error(const mcl_engine* h_or_null);
/* The engine's CUDA stream (cudaStream_t) so a caller can order its own work / events against it. */
void* mcl_stream(mcl_engine* h);
int  mcl_sync(mcl_engine* h);

/* ---- multi-GPU: one engine per process per GPU; rank r scores the global particle slice [r*N/R, (r+1)*N/R) ------ */
/* nccl_unique_id: the 128-byte ncclUniqueId made by mcl_comm_unique_id on rank 0 and shipped by the host
 * (torch.distributed / MPI / a socket).  The map is replicated: call mcl_set_map on every rank. */
int  mcl_comm_unique_id(void* id128_out);
int  mcl_comm_init(mcl_engine* h, const void* nccl_unique_id128, int rank, int world_size);

/* ---- OccupancyGrid device mirror (occupancy_grid.cpp:63-71 read side; slam.cpp:274-281 mutates it every update) -- */
/* cells: row-major int8, index y*width+x (occupancy_grid.hpp:208).  cells_per_meter is explicit because the
 * reference's can be stale after loadFromFile (occupancy_grid.cpp:151-159). */
int  mcl_set_map(mcl_engine* h, const int8_t* cells, int width, int height, float origin_x, float origin_y,
                 float meters_per_cell, float cells_per_meter);
int  mcl_update_map_rect(mcl_engine* h, int x0, int y0, int w, int hgt, const int8_t* src, int src_stride);
int  mcl_read_map_rect(mcl_engine* h, int x0, int y0, int w, int hgt, int8_t* dst, int dst_stride);
/* Mapping::updateMap (mapping.cpp:17-127) applied to the device mirror: +hit_odds at the endpoint cell of every ray of
 * the MovingLaserScan between previous_pose and pose (ranges <= max_laser_distance), then -miss_odds along each ray's
 * Bresenham walk, both saturating in int8; bit-identical to the reference's sequential loops (saturating adds of one
 * sign commute).  initialized = Mapping::initialized_: 0 reproduces the reference's first call, which changes no cell.
 * rect_xywh_out (4 ints, may be NULL): the rectangle of cells that may have changed, for mcl_read_map_rect. */
int  mcl_map_update(mcl_engine* h, const mcl_pose_t* previous_pose, const mcl_pose_t* pose, int initialized,
                    const float* ranges, const float* thetas, const int64_t* times, int num_ranges,
                    float max_laser_distance, int hit_odds, int miss_odds, int* rect_xywh_out);

/* ObstacleDistanceGrid::setDistances (planning/obstacle_distance_grid.cpp:73-188) of the device mirror: out[y*width+x] =
 * the reference's float distance of cell (x, y) -- 0 for cells with log-odds >= 0, else 0.1f accumulated once per
 * four-connected step to the nearest such cell, -1 where the brushfire never arrives.  Bit-identical to the reference. */
int  mcl_distance_grid(mcl_engine* h, float* out);

/* ---- particle state -------------------------------------------------------------------------------------------- */
/* ParticleFilter::initializeFilterAtPose (particle_filter.cpp:16-34) with the intended weight 1.0/N and a seeded
 * counter-based generator instead of std::random_device. */
int  mcl_init_at_pose(mcl_engine* h, float x, float y, float theta, int64_t utime, uint64_t seed);
/* Extension for global localisation (no reference equivalent): x,y uniform over the map, theta uniform in [-pi,pi).
 * The sample is stratified: the map is cut into equal blocks (about 1024 particles each), consecutive particles fill
 * one block after the other in serpentine order, each uniform inside its block -- so consecutive particles are spatial
 * neighbours, which systematic resampling preserves and the sensor kernels exploit (one map window per batch). */
int  mcl_init_uniform(mcl_engine* h, int64_t utime, uint64_t seed);
/* AoS particle_t in/out (ParticleFilter::particles, particle_filter.cpp:75-81).  All particles must share one
 * pose.utime and one parent_pose.utime (true for every cloud the reference produces).  Export writes
 * min(max_n, ceil(N/stride)) particles: every stride-th one. */
int  mcl_import_particles(mcl_engine* h, const mcl_particle_t* aos, int64_t n);
/* With mcl_comm_init'd engines export is COLLECTIVE (every rank calls it): parent poses are exchanged first. */
int  mcl_export_particles(mcl_engine* h, mcl_particle_t* aos, int64_t max_n, int64_t stride, int64_t* n_out);

/* A WEIGHTED sub-sample for SLAM_PARTICLES consumers (botgui draws the cloud from stack arrays, drawing_functions.cpp:
 * 123-125, so it cannot take millions): `count` particles drawn by systematic sampling over the normalised weights
 * (the filter's own resampling rule with `count` draws, offset u01 in [0, 1)), each exported with weight 1/count.
 * The weights must sum to 1 (after mcl_init_*, mcl_update or mcl_normalize).  Single-GPU engines. */
int  mcl_export_weighted(mcl_engine* h, mcl_particle_t* aos, int64_t count, double u01, int64_t* n_out);

/* ---- host-side scalar part of the action model: ActionModel::updateAction (action_model.cpp:22-75) ------------- */
void mcl_action_reset(mcl_action_t* a);
int  mcl_action_update(mcl_action_t* a, const mcl_pose_t* odometry);   /* returns moved (1/0) */

/* ---- stages, each usable alone for stage-wise parity ------------------------------------------------------------ */
/* ParticleFilter::resamplePosteriorDistribution (particle_filter.cpp:84-103).  r = rand()/RAND_MAX/N is injected.
 * weights_or_null: N doubles replacing the particles' weights first.  The running sum c reproduces the reference's
 * sequential double rounding exactly, so indices are bit-identical to the reference loop.  indices_out_or_null gets
 * the N source indices.  Particles (pose, parent_pose, weight) are gathered in place like prior[m] = posterior_[i]. */
int  mcl_resample(mcl_engine* h, double r, const double* weights_or_null, int32_t* indices_out_or_null);
/* ActionModel::applyAction over all particles (action_model.cpp:78-103, particle_filter.cpp:106-113).
 * noise3n_or_null: N x (rot1, trans, rot2) float draws to inject (the reference's recorded draws); NULL = draw from the
 * engine's Philox4x32-10 stream keyed by (seed, update counter, global particle index). */
int  mcl_apply_action(mcl_engine* h, const mcl_action_t* a, int64_t utime, const float* noise3n_or_null);
/* SensorModel::likelihood for every particle (sensor_model.cpp:14-86, moving_laser_scan.cpp:8-39,
 * interpolation.hpp:24-50).  Scan = lidar_t's arrays (lcmtypes/lidar_t.lcm).  scores_out_or_null: N doubles. */
int  mcl_score(mcl_engine* h, const float* ranges, const float* thetas, const int64_t* times, int num_ranges,
               double* scores_out_or_null);
/* ParticleFilter::computeNormalizedPosterior's floor / sum / divide (particle_filter.cpp:120-138), on the scores of
 * the last mcl_score.  The sum has the reference's sequential-double semantics.  weights_out_or_null: N doubles. */
int  mcl_normalize(mcl_engine* h, double* weights_out_or_null);
/* ParticleFilter::estimatePosteriorPose (particle_filter.cpp:144-160): weighted mean x,y and circular mean theta. */
int  mcl_estimate(mcl_engine* h, mcl_pose_t* pose_out);

/* ---- fused update: ParticleFilter::updateFilter (particle_filter.cpp:37-52) -------------------------------------- */
/* a: filled by mcl_action_update for this odometry.  If !a->moved nothing runs and pose_out is the previous estimate
 * with utime = odometry_utime.  r: the resample draw.  noise3n_or_null as in mcl_apply_action. */
int  mcl_update(mcl_engine* h, const mcl_action_t* a, int64_t odometry_utime, const float* ranges, const float* thetas,
                const int64_t* times, int num_ranges, double r, const float* noise3n_or_null, mcl_pose_t* pose_out);
/* ParticleFilter::updateFilterActionOnly (particle_filter.cpp:54-65): action model only, no resample, no scoring. */
int  mcl_update_action_only(mcl_engine* h, const mcl_action_t* a, int64_t odometry_utime,
                            const float* noise3n_or_null);

/* Device-resident pipelining for throughput runs: upload a scan once, then enqueue updates without host round trips.
 * mcl_update_enqueue neither copies inputs nor synchronises; the estimate is read back later with mcl_read_estimate
 * (which synchronises).  r_or_negative < 0 draws r from the Philox stream. */
int  mcl_upload_scan(mcl_engine* h, const float* ranges, const float* thetas, const int64_t* times, int num_ranges,
                     int64_t odometry_utime);
int  mcl_update_enqueue(mcl_engine* h, const mcl_action_t* a, int64_t odometry_utime, double r_or_negative);
int  mcl_read_estimate(mcl_engine* h, mcl_pose_t* pose_out);

/* ---- introspection ------------------------------------------------------------------------------------------------ */
int  mcl_get_stats(mcl_engine* h, mcl_stats* out);
int  mcl_set_gather_counting(mcl_engine* h, int on);   /* debug counter of map reads (slower scoring kernel) */
/* Roofline denominator measured on this device: uniformly random 1-byte reads over a footprint of map_bytes, L1
 * bypassed, one per lane.  Returns sectors (32 B) per second. */
int  mcl_measure_gather_peak(mcl_engine* h, int64_t footprint_bytes, int64_t reads, double* sectors_per_s_out);
/* Digest of this rank's slice of the last update: four 64-bit sums (mod 2^64) of mixed (global particle index, bit
 * pattern) pairs of the resample indices, the half-unit scores, the normalised weights and the poses.  The sum over the
 * ranks is independent of the GPU count iff the clouds are bit-identical (bench.py prints it; the multi-GPU parity test
 * compares whole clouds). */
int  mcl_debug_digest(mcl_engine* h, uint64_t* digest4_out);
/* glibc-sincosf restatement evaluated on the device for n floats (test hook for the trig parity contract). */
int  mcl_debug_sincosf(mcl_engine* h, const float* x, int64_t n, float* sin_out, float* cos_out);
/* Largest absolute error of the SFU sine / cosine the certified float pass uses, over EVERY float in [lo, hi] against
 * double-precision sin/cos (test hook: the certification's error budget assumes a bound on it). */
int  mcl_debug_fast_trig_error(mcl_engine* h, float lo, float hi, double* max_sin_err, double* max_cos_err);
/* Certification margin probe for the current map, particles and scan: the largest deviation (cells) between the float
 * pass's endpoint / extended point and the reference's exactly-rounded ones, and the eps the certification would assume
 * for a whole-grid window (0 when the float pass is not applicable).  Test hook: deviations must stay below eps. */
int  mcl_debug_fast_margin(mcl_engine* h, double* max_dev_endpoint, double* max_dev_extended, double* eps_out);

#ifdef __cplusplus
}
#endif
#endif /* MCL_CUDA_H */
