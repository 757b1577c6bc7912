// libmcl_cuda.so -- engine object and the C ABI declared in include/mcl_cuda.h.
// One engine = one GPU = one stream.  All particle state is SoA in HBM and never leaves the device on the fast path;
// per update the host sends ~6 KB (prepared beams + scalars) and reads back one pose.
#include "../../include/mcl_cuda.h"
#include "mcl_kernels.cuh"
#include "mcl_table.cuh"
#include "mcl_shard.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#ifdef MCL_WITH_NCCL
#include <nccl.h>
#endif

using namespace mcl;

namespace {

thread_local std::string g_last_error;

struct PoseSoA {
    float *x = nullptr, *y = nullptr, *th = nullptr;
};

}  // namespace

// What a blocking call reads back (read_back()): counters[0..5] = overruns, gathers, fall-back chunks, (double) weight sum,
// (double) sum w^2, deferred evaluations; counters[6] = barrier time-out flag; est = the pose estimate.
struct Readback { unsigned long long counters[7]; float est[4]; int culled; int pad; };
struct mcl_engine {
    mcl_params params;
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // partition
    long long n = 0;            // global particle count
    int rank = 0, world = 1;
    long long lo = 0, hi = 0;   // slice scored / moved by this rank
#ifdef MCL_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
    // peer push of pose slices over NVLink by the copy engines (no SMs), overlapped with the sensor kernel
    std::vector<float*> peer_pose_block;         // IPC-mapped pose_block of every rank (own entry = pose_block)
    // exchange block (mcl_shard.cuh): weights, sequential-sum maps, estimate partials, barrier flags; IPC-mapped by
    // every rank so that slices are exchanged by peer stores / loads
    unsigned char* xblock = nullptr;
    XLayout xl{};
    XPeers xp{};
    bool xpeer = false;                          // peers' exchange blocks are mapped (multi-GPU sharded stages)
    int xepoch[kXFlagSlots] = {0};
    int* fb_count = nullptr;                     // raw-chunk side-buffer slots handed out by this rank (per stage)
    long long* xrange = nullptr;                 // [2] groups this rank's children draw from
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_action_done = nullptr, ev_push_done = nullptr;
    bool peer_push = false;
    bool push_pending = false;
    int* barrier_word = nullptr;

    // particle state (global-sized arrays; every rank holds the full cloud, only [lo,hi) is computed locally)
    PoseSoA pose[2], parent[2];
    float* pose_block = nullptr;                 // one allocation behind pose[0..1].{x,y,th}: [2][3][N] floats
    int cur = 0;
    double* weight[2] = {nullptr, nullptr};      // inside xblock
    int wcur = 0;
    int32_t* score_block = nullptr;     // [N] half-unit scores (only the rank's slice is computed)
    int32_t* score2 = nullptr;
    int32_t* idx = nullptr;
    long long pose_utime = 0, parent_utime = 0;
    bool have_particles = false, have_scores = false;

    // sequential-sum workspace
    long long n1 = 0, n2 = 0;
    double *cum = nullptr, *sums = nullptr, *cin1 = nullptr, *cin2 = nullptr, *total = nullptr;
    double *tile_sums = nullptr, *tile_excl = nullptr;
    int *ebias = nullptr, *gebias = nullptr, *opened = nullptr;
    long long *q0 = nullptr, *q1 = nullptr, *g0 = nullptr, *g1 = nullptr, *fallbacks = nullptr;
    unsigned long long* overruns = nullptr;
    unsigned long long* gather_counter = nullptr;
    unsigned long long* deferred_counter = nullptr;
    Beam* map_beams = nullptr;          // map update: its own prepared beams (device / pinned), counts window, flag
    Beam* map_beams_host = nullptr;
    int map_beams_cap = 0;
    uint32_t* map_counts = nullptr;
    size_t map_counts_cap = 0;
    int* map_flag = nullptr;
    int* map_flag_host = nullptr;
    BatchWindow* windows = nullptr;     // per-batch map windows (global localisation)
    size_t windows_cap = 0;
    uint32_t* masks = nullptr;          // two-pass sensor path: uncertain-beam bits, [word][virtual lane]
    size_t masks_bytes = 0;
    double* ess_acc = nullptr;
    int* bbox = nullptr;
    // table sensor path (mcl_table.cuh): device-side window plan, its pinned copy (the host's hint for the next update)
    TabPlan* tab_plan = nullptr;
    int* tab_box = nullptr;             // bounding box of the cloud (ordered-int atomics); table_plan_kernel re-arms it
    int* tab_build = nullptr;           // [0] table entries needed, [1] CTAs whose table overflowed
    struct TabHint { TabPlan plan; int build[8]; };
    TabHint* tab_hint = nullptr;        // pinned
    TabHint* tab_hint_dev = nullptr;    // device: tab_plan and tab_build point into it
    cudaEvent_t ev_tab_hint = nullptr;
    bool tab_hint_pending = false;
    bool tab_ok = false;                // the last plan the host has seen allows the table pass
    int tab_blocked = 0;                // updates to keep off the table pass after an overflow
    int tab_variant = 0;                // kTabSingle16 .. kTabBatch8 (mcl_table.cuh): what the next pass launches
    int tab_excluded = 0;               // variants whose score table overflowed on this cloud
    uint8_t* tab_cull = nullptr;        // per beam: 1 = the table pass skips it (table_cull_kernel)
    std::unordered_map<const void*, int> smem_attr;     // kernels whose dynamic shared-memory limit has been raised, to what
    int cull_interval = 1, cull_wait = 0;   // the culling pass runs every cull_interval-th update while it finds nothing
    int tab_batch = kTabBatch;          // kTabBatch = 4096, halved down to kTabBatchSmall = 1024 until the windows fit
    int4* tab_bboxes = nullptr;         // bounding box per batch
    size_t tab_bboxes_cap = 0;
    bool scan_finite = true;

    // estimate
    int est_count = 0;
    float* est_out = nullptr;          // device float4
    float* est_host = nullptr;         // pinned float4
    mcl_pose_t last_estimate{};

    // map mirror
    int8_t* map = nullptr;
    int8_t* map_fast = nullptr;         // the fast pass's derived view of the map (derive_fast_map_kernel), same pitch
    int8_t* map_lf = nullptr;           // likelihood field (sensor_mode 1), same pitch; rebuilt lazily after map changes
    uint8_t* map_cls = nullptr;         // class map of the score-table pass (mcl_table.cuh: derive_class_map_kernel): the grid
    unsigned long long* map_pack = nullptr;     // plus a four-cell apron, cpitch bytes per row; score-table entries of its cells
    int cpitch = 0;
    uint16_t* dt_steps = nullptr;       // distance-transform scratch: [H][W] steps
    bool lf_dirty = true;
    DevGrid grid{};
    float meters_per_cell = 0.05f;
    bool have_map = false;

    // scan
    Beam* beams = nullptr;             // device
    Beam* beams_host = nullptr;        // pinned
    int beams_cap = 0, num_beams = 0;
    float max_range = 0.0f;
    float max_abs_theta = 0.0f;
    double ratio_lo = 0.0, ratio_hi = 1.0;
    bool scan_interp = false;
    bool have_scan = false;

    // staging
    float* noise = nullptr;            // device, 3 floats per particle, allocated on first injection
    void* staging = nullptr;           // device scratch for AoS / double vectors
    size_t staging_bytes = 0;
    int* host_bbox = nullptr;          // pinned 4 ints
    int* host_bbox_init = nullptr;     // pinned: the empty box
    Readback* readback_dev = nullptr;             // see read_back()
    Readback* readback_host = nullptr;            // pinned

    // stats
    mcl_stats stats{};
    mutable double stats_eps = 0.0;
    bool count_gathers = false;
    uint64_t seed = 0x5eedULL;
    uint32_t update_no = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // MCL_PROFILE=1: CUDA events around the parts of an update (blocking mcl_update only), averaged and printed to
    // stderr by mcl_destroy
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<const char*> prof_label;
    size_t prof_n = 0;
    std::vector<double> prof_sum;
    long prof_updates = 0;
    int launches = 0;
    int collectives = 0;
};

namespace {

void prof_mark(mcl_engine* h, const char* label)
{
    if (!h->prof) return;
    if (h->prof_n == h->prof_ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->prof_ev.push_back(e);
        h->prof_label.push_back(label);
        h->prof_sum.push_back(0.0);
    }
    h->prof_label[h->prof_n] = label;
    cudaEventRecord(h->prof_ev[h->prof_n++], h->stream);
}

void prof_collect(mcl_engine* h)          // after a stream synchronisation
{
    if (!h->prof) return;
    for (size_t i = 1; i < h->prof_n; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->prof_ev[i - 1], h->prof_ev[i]);
        h->prof_sum[i] += ms;
    }
    ++h->prof_updates;
    h->prof_n = 0;
}

int fail(mcl_engine* h, int code, const char* fmt, ...);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and size (the call costs the host microseconds)
template <class K>
int raise_smem(mcl_engine* h, K kernel, size_t bytes)
{
    auto it = h->smem_attr.find((const void*)kernel);
    if (it != h->smem_attr.end() && it->second >= (int)bytes) return 0;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(h, MCL_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    h->smem_attr[(const void*)kernel] = (int)bytes;
    return 0;
}

int fail(mcl_engine* h, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (h) h->err = buf;
    return code;
}

#define CK(call)                                                                                                     \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return fail(h, MCL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define CKL(h)                                                                                                       \
    do {                                                                                                             \
        cudaError_t e__ = cudaGetLastError();                                                                        \
        if (e__ != cudaSuccess)                                                                                      \
            return fail(h, MCL_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__,      \
                        __LINE__);                                                                                   \
        ++(h)->launches;                                                                                             \
    } while (0)

template <class T>
int dev_alloc(mcl_engine* h, T** p, size_t count)
{
    CK(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return MCL_OK;
}

int grid_for(const mcl_engine* h, long long work, int block, int per_sm = 8)
{
    long long want = (work + block - 1) / block;
    long long cap = (long long)h->sm_count * per_sm;
    return (int)std::max<long long>(1, std::min(want, cap));
}

int ensure_staging(mcl_engine* h, size_t bytes)
{
    if (bytes <= h->staging_bytes) return MCL_OK;
    if (h->staging) cudaFree(h->staging);
    h->staging = nullptr;
    h->staging_bytes = 0;
    CK(cudaMalloc(&h->staging, bytes));
    h->staging_bytes = bytes;
    return MCL_OK;
}

// ---- cross-rank barrier on the engine's stream: flags in the exchange blocks (mcl_shard.cuh) ------------------------------
// Everything this rank enqueued before the call (peer stores included) is complete before any rank gets past it.
int xbarrier(mcl_engine* h, int slot)
{
    if (h->world == 1) return MCL_OK;
    if (!h->xpeer) return fail(h, MCL_ERR_COMM, "exchange blocks are not mapped");
    const int epoch = ++h->xepoch[slot];
    xbarrier_kernel<<<1, 32, 0, h->stream>>>(h->xp, h->xl.flags, h->xl.err, slot, epoch);
    CKL(h);
    ++h->collectives;
    return MCL_OK;
}

// ---- exact sequential running sum of the weights (buffer wbuf of the exchange block), sharded over the ranks ----------------
// Leaves the exact total in h->total and the exact entry sum of every group in h->cin2 on every rank.
int seq_total(mcl_engine* h, int wbuf, double* zero_after = nullptr)
{
    const long long n = h->n, n1 = h->n1, n2 = h->n2;
    const double* w = h->weight[wbuf];
    const long long tile_lo = h->lo / (kL1 * kSeqTileChunks);
    const long long tile_hi = (h->hi + kL1 * kSeqTileChunks - 1) / (kL1 * kSeqTileChunks);
    const long long group_lo = h->lo / kSliceAlign, group_hi = (h->hi + kSliceAlign - 1) / kSliceAlign;
    const int tiles = (int)(tile_hi - tile_lo);
    prof_mark(h, "seq:begin");
    if (tiles > 0) {
        xseq_chunk_sums_kernel<<<tiles, 128, 0, h->stream>>>(w, n, n1, tile_lo, h->sums, h->tile_sums);
        CKL(h);
    }
    xseq_tile_scan_kernel<<<1, 1024, 0, h->stream>>>(h->tile_sums, tile_lo, tile_hi, h->tile_excl, h->xp, h->xl.tot, h->fb_count);
    CKL(h);
    prof_mark(h, "seq:S1+S2");
    int rc = xbarrier(h, 0);
    if (rc) return rc;
    prof_mark(h, "seq:barrierA");
    if (tiles > 0) {
        XSeqOut o{h->xl.q0, h->xl.q1, h->xl.eb, h->xl.fbraw};
        xseq_chunk_maps_kernel<<<tiles, 128, 0, h->stream>>>(w, n, n1, tile_lo, h->sums, h->tile_excl,
                                                            (const double*)(h->xblock + h->xl.tot), h->fb_count, h->xp, o);
        CKL(h);
        xseq_group_maps_kernel<<<(int)((group_hi - group_lo + 3) / 4), 128, 0, h->stream>>>(
            h->ebias, h->q0, h->q1, n1, group_lo, group_hi, h->xp, h->xl.g0, h->xl.g1, h->xl.ge);
        CKL(h);
    }
    prof_mark(h, "seq:S3+S4");
    rc = xbarrier(h, 1);
    if (rc) return rc;
    prof_mark(h, "seq:barrierB");
    {
        const size_t walk_smem = xseq_walk_smem(n2);                 // group maps + the pre-staged chunk maps and raw chunks
        const int staged = walk_smem + 4096 <= (size_t)h->max_smem_optin ? 1 : 0;
        if (staged) { int rc2 = raise_smem(h, xseq_walk_kernel, walk_smem); if (rc2) return rc2; }
        xseq_walk_kernel<<<1, staged ? 1024 : 32, staged ? walk_smem : 0, h->stream>>>(
            n, n1, n2, h->ebias, h->q0, h->q1, h->gebias, h->g0, h->g1, h->cin2, h->cin1, h->opened, h->total,
            h->fallbacks, staged, h->xp, h->xl.w[wbuf], h->xl.fbraw, zero_after);
        CKL(h);
    }
    prof_mark(h, "seq:walk");
    return MCL_OK;
}

// ---- multi-GPU slice exchange: every rank contributes buf[lo_r, hi_r) and ends with the complete array ------------------
// In place on the engine's stream (no host synchronisation).  Equal slices use ncclAllGather; ragged ones a group of
// broadcasts.  These are the only collectives of an update: 3 x 4 B/particle of poses and 4 B/particle of scores.
int exchange_slices(mcl_engine* h, void* buf, size_t elem)
{
    if (h->world == 1) return MCL_OK;
#ifdef MCL_WITH_NCCL
    char* base = (char*)buf;
    ncclResult_t rc;
    {
        ncclGroupStart();
        rc = ncclSuccess;
        for (int r = 0; r < h->world && rc == ncclSuccess; ++r) {
            const long long lo = h->xp.lo[r], hi = h->xp.lo[r + 1];
            if (hi <= lo) continue;
            rc = ncclBroadcast(base + (size_t)lo * elem, base + (size_t)lo * elem, (size_t)(hi - lo) * elem, ncclChar, r,
                               h->comm, h->stream);
        }
        ncclResult_t rc2 = ncclGroupEnd();
        if (rc == ncclSuccess) rc = rc2;
    }
    if (rc != ncclSuccess) return fail(h, MCL_ERR_COMM, "NCCL slice exchange failed: %s", ncclGetErrorString(rc));
    ++h->collectives;
    return MCL_OK;
#else
    (void)buf; (void)elem;
    return fail(h, MCL_ERR_COMM, "library built without NCCL");
#endif
}

// Cross-rank barrier on the engine's stream (a 4-byte all-reduce), used where no other collective follows a push.
int rank_barrier(mcl_engine* h)
{
    if (h->world == 1) return MCL_OK;
#ifdef MCL_WITH_NCCL
    if (ncclAllReduce(h->barrier_word, h->barrier_word, 1, ncclInt, ncclMax, h->comm, h->stream) != ncclSuccess)
        return fail(h, MCL_ERR_COMM, "NCCL barrier failed");
    return MCL_OK;
#else
    return fail(h, MCL_ERR_COMM, "library built without NCCL");
#endif
}

// Orders the engine's stream after this rank's outstanding pose pushes.
int join_pushes(mcl_engine* h)
{
    if (h->push_pending) {
        CK(cudaStreamWaitEvent(h->stream, h->ev_push_done, 0));
        h->push_pending = false;
    }
    return MCL_OK;
}

// Refreshes the fast pass's derived map over the rectangle [x0, x0+w) x [y0, y0+hgt) widened by the 2-cell look-ahead.
// Classes (and score-table entries) of the cells whose 5 x 5 neighbourhood meets the changed rectangle, apron included.
int refresh_class_map(mcl_engine* h, int x0, int y0, int w, int hgt)
{
    const bool lf = h->params.sensor_mode == 1;
    const int xa = std::max(-kTabApron, x0 - 2), ya = std::max(-kTabApron, y0 - 2);
    const int xb = std::min(h->grid.width + kTabApron, x0 + w + 2), yb = std::min(h->grid.height + kTabApron, y0 + hgt + 2);
    if (xb <= xa || yb <= ya) return MCL_OK;
    TabArgs a{};
    a.grid = h->grid;
    a.fast_cells = h->map_fast;
    a.lf_cells = lf ? h->map_lf : nullptr;
    const long long total = (long long)(xb - xa) * (yb - ya);
    derive_class_map_kernel<<<grid_for(h, total, 256), 256, 0, h->stream>>>(a, h->map_cls, h->map_pack, h->cpitch, xa, ya, xb - xa,
                                                                           yb - ya);
    CKL(h);
    return MCL_OK;
}

int refresh_fast_map(mcl_engine* h, int x0, int y0, int w, int hgt)
{
    h->lf_dirty = true;
    const int xa = std::max(0, x0 - 2), ya = std::max(0, y0 - 2);
    const int xb = std::min(h->grid.width, x0 + w + 2), yb = std::min(h->grid.height, y0 + hgt + 2);
    if (xb <= xa || yb <= ya) return MCL_OK;
    const long long total = (long long)(xb - xa) * (yb - ya);
    derive_fast_map_kernel<<<grid_for(h, total, 256), 256, 0, h->stream>>>(h->map, h->map_fast, h->grid.width, h->grid.height,
                                                                          h->grid.pitch, xa, ya, xb - xa, yb - ya);
    CKL(h);
    // (likelihood-field mode derives its classes from the field, which is rebuilt lazily: refresh_likelihood_field)
    if (h->params.sensor_mode != 1) return refresh_class_map(h, xa - 2, ya - 2, xb - xa + 4, yb - ya + 4);
    return MCL_OK;
}

// ---- scan preparation (host): valid-beam compaction + interpolation ratios ------------------
int prepare_scan(mcl_engine* h, const float* ranges, const float* thetas, const int64_t* times, int nb,
                 long long t_begin, long long t_end)
{
    if (nb < 0 || (nb > 0 && (!ranges || !thetas || !times))) return fail(h, MCL_ERR_INVALID, "bad scan arrays");
    if (nb > h->beams_cap) {
        if (h->beams) cudaFree(h->beams);
        if (h->beams_host) cudaFreeHost(h->beams_host);
        h->beams = nullptr; h->beams_host = nullptr;
        const int cap = std::max(nb, 1024);
        CK(cudaMalloc((void**)&h->beams, sizeof(Beam) * cap));
        CK(cudaMallocHost((void**)&h->beams_host, sizeof(Beam) * cap));
        h->beams_cap = cap;
    }
    // interpolation.hpp:29-36: equal utimes -> every ray from the end pose; else ratio in double, unclamped.
    const bool interp = t_begin != t_end;
    const double denom = (double)(t_end - t_begin);
    int k = 0;
    float mx = 0.0f, mth = 0.0f;
    double rlo = 1.0, rhi = 1.0;
    bool finite = true, all_finite = true;
    for (int i = 0; i < nb; ++i) {
        if (ranges[i] > h->params.min_range) {                 // moving_laser_scan.cpp:24
            Beam b;
            b.range = ranges[i];
            b.theta = thetas[i];
            b.ratio = interp ? (double)(times[i] - t_begin) / denom : 1.0;
            h->beams_host[k++] = b;
            if (std::isfinite(ranges[i])) mx = std::max(mx, ranges[i]);
            else all_finite = false;
            finite = finite && std::isfinite(thetas[i]);
            mth = std::max(mth, std::fabs(thetas[i]));
            rlo = (k == 1) ? b.ratio : std::min(rlo, b.ratio);
            rhi = (k == 1) ? b.ratio : std::max(rhi, b.ratio);
        }
    }
    h->num_beams = k;
    h->max_range = mx;
    h->max_abs_theta = finite ? mth : INFINITY;
    h->scan_finite = finite && all_finite;
    h->ratio_lo = rlo; h->ratio_hi = rhi;
    h->scan_interp = interp;
    if (k > 0) CK(cudaMemcpyAsync(h->beams, h->beams_host, sizeof(Beam) * k, cudaMemcpyHostToDevice, h->stream));
    h->have_scan = true;
    h->stats.valid_beams = k;
    return MCL_OK;
}

// ---- sensor-model launch ---------------------------------------------------------------------------------------------
// Persistent CTAs: exactly one resident wave (SMs x CTAs that really fit: registers and shared memory).
template <class K>
int launch_persistent(mcl_engine* h, K kernel, const ScoreArgs& a, int threads, size_t smem, long long want_blocks)
{
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    per_sm = std::max(per_sm, 1);
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want_blocks, (long long)h->sm_count * per_sm));
    kernel<<<blocks, threads, smem, h->stream>>>(a);
    CKL(h);
    return MCL_OK;
}

// which: 0 = literal restatement for every beam, 1 = certified float pass, 2 = exact pass over the deferred beams
template <int G, bool INTERP, bool TILE>
int launch_score_g(mcl_engine* h, const ScoreArgs& a, int which, size_t smem, long long local)
{
    const bool cnt = h->count_gathers;
    if (which == 1) {
        const long long blocks = (local + MCL_FAST_THREADS / G - 1) / (MCL_FAST_THREADS / G);
        return cnt ? launch_persistent(h, score_fast_kernel<G, INTERP, TILE, true, false>, a, MCL_FAST_THREADS, smem, blocks)
                   : launch_persistent(h, score_fast_kernel<G, INTERP, TILE, false, false>, a, MCL_FAST_THREADS, smem, blocks);
    }
    if (which == 2) {
        const long long blocks = (local * G + MCL_DEF_THREADS - 1) / MCL_DEF_THREADS;
        return cnt ? launch_persistent(h, score_deferred_kernel<G, INTERP, TILE, true, false>, a, MCL_DEF_THREADS, smem, blocks)
                   : launch_persistent(h, score_deferred_kernel<G, INTERP, TILE, false, false>, a, MCL_DEF_THREADS, smem, blocks);
    }
    const long long blocks = (local + MCL_SCORE_THREADS / G - 1) / (MCL_SCORE_THREADS / G);
    return cnt ? launch_persistent(h, score_kernel<G, INTERP, TILE, true, false>, a, MCL_SCORE_THREADS, smem, blocks)
               : launch_persistent(h, score_kernel<G, INTERP, TILE, false, false>, a, MCL_SCORE_THREADS, smem, blocks);
}

template <bool INTERP, bool TILE>
int launch_score_it(mcl_engine* h, int G, const ScoreArgs& a, int which, size_t smem, long long local)
{
    switch (G) {
        case 1: return launch_score_g<1, INTERP, TILE>(h, a, which, smem, local);
        case 2: return launch_score_g<2, INTERP, TILE>(h, a, which, smem, local);
        case 4: return launch_score_g<4, INTERP, TILE>(h, a, which, smem, local);
        case 8: return launch_score_g<8, INTERP, TILE>(h, a, which, smem, local);
        case 16: return launch_score_g<16, INTERP, TILE>(h, a, which, smem, local);
        default: return launch_score_g<32, INTERP, TILE>(h, a, which, smem, local);
    }
}

// per-batch windows: one lane per particle, shared-memory tiles, fixed CTA shapes (mcl_kernels.cuh: kBatch*)
template <bool INTERP>
int launch_score_batch(mcl_engine* h, const ScoreArgs& a, int which, size_t smem)
{
    const bool cnt = h->count_gathers;
    const long long blocks = a.num_batches;
    if (which == 1)
        return cnt ? launch_persistent(h, score_fast_kernel<1, INTERP, true, true, true>, a, kBatchFastThreads, smem, blocks)
                   : launch_persistent(h, score_fast_kernel<1, INTERP, true, false, true>, a, kBatchFastThreads, smem, blocks);
    if (which == 2)
        return cnt ? launch_persistent(h, score_deferred_kernel<1, INTERP, true, true, true>, a, kBatchDefThreads, smem, blocks)
                   : launch_persistent(h, score_deferred_kernel<1, INTERP, true, false, true>, a, kBatchDefThreads, smem, blocks);
    return cnt ? launch_persistent(h, score_kernel<1, INTERP, true, true, true>, a, kBatchExactThreads, smem, blocks)
               : launch_persistent(h, score_kernel<1, INTERP, true, false, true>, a, kBatchExactThreads, smem, blocks);
}

int launch_score(mcl_engine* h, int G, bool tile, bool batch, const ScoreArgs& a, int which, size_t smem, long long local)
{
    if (batch) return h->scan_interp ? launch_score_batch<true>(h, a, which, smem) : launch_score_batch<false>(h, a, which, smem);
    if (h->scan_interp)
        return tile ? launch_score_it<true, true>(h, G, a, which, smem, local) : launch_score_it<true, false>(h, G, a, which, smem, local);
    return tile ? launch_score_it<false, true>(h, G, a, which, smem, local) : launch_score_it<false, false>(h, G, a, which, smem, local);
}

// Error budget of the certified fast pass (mcl_device.cuh: score_beam_fast) for a window [x0, x0+w) x [y0, y0+hh) of
// global cells.  All terms in cells; u = 2^-24 is the float unit roundoff; every bound is for |coordinates| <= Cm,
// |world metres| <= Xm, ray length <= Rc cells, |dS| <= max_shift, |rho| <= rho_max, |theta_beam| <= 6.3.
//   reference vs the real-valued model:  ox rounding (x cpm) + sx rounding + endpoint-add rounding + two product
//     roundings + Rc x (angle roundings: theta_r, the subtraction, two wrap steps = 20u; glibc sincosf <= 1 ulp)
//   fast pass vs the same model:  S_b rounding + fma rounding + endpoint-add rounding + dS and rho roundings + rc and
//     product roundings + Rc x (angle roundings 32u + measured SFU error kFastTrigErr)
// eps = 1.25 x the sum.  A coordinate is certain when it is further than eps + 2^-11 (fixed-point rounding) from an
// integer; a direction when both octant discriminants exceed 3(1 + eps).
// max_dim: the largest window width/height the plan has to cover (the window's own for a single window, the largest
// over the batches for per-batch windows).
FastPlan fast_plan(const mcl_engine* h, long long x0, long long y0, long long w, long long hh, long long pitch,
                   long long max_dim)
{
    FastPlan fp{};
    const double cpm = h->grid.cells_per_meter;
    const double Rc = (double)h->max_range * cpm;
    const double rho_max = std::max(std::fabs(h->ratio_lo), std::fabs(h->ratio_hi));
    const double Cm = (double)std::max(x0 + w, y0 + hh) + 1.0;
    const double We = (double)max_dim + Rc + 8.0;     // magnitude of the window-relative coordinates the budget covers
    // min_range * cpm >= 2.5: the endpoint is never the robot's own cell (the reference's zero-difference step rule
    // is then out of play); pitch < 2^20: the float-assembled step offset is exact
    const bool usable = h->params.sensor_path != 1 && h->num_beams > 0 && std::isfinite(Rc) && cpm > 0.0 &&
                        (double)h->params.min_range * cpm >= 2.5 && h->max_abs_theta <= 6.3f && h->ratio_lo >= -1.0 &&
                        h->ratio_hi <= 2.0 && w >= 3 && hh >= 3 && Cm <= 4090.0 && x0 >= -4 && y0 >= -4 &&
                        pitch < (1 << 20) && max_dim + 8 < 4096 &&
                        h->num_beams <= 2048;     // score_deferred_kernel's queue entries hold 6 bits of mask-word index
    if (!usable) return fp;
    // v + 1.5*2^(23-FB) must stay in one binade for every coordinate whose bits are used: those inside the window
    const int fb = max_dim + 8 < 1024 ? 12 : (max_dim + 8 < 2048 ? 11 : 10);
    const double u = 5.9604644775390625e-08;
    const double Xm = Cm / cpm + std::max(std::fabs((double)h->grid.origin_x), std::fabs((double)h->grid.origin_y));
    const double max_shift = 64.0;
    const double Ce = Cm + Rc;            // endpoints beyond the grid (certified as "outside") reach this far
    const double e_ref = cpm * u * Xm + 2.0 * u * Ce + 2.0 * u * Rc + Rc * (20.0 * u + 1.2e-7) + 1e-9;
    const double e_apx = 3.0 * u * We + (1.0 + 2.0 * rho_max) * u * max_shift + 2.0 * u * Rc +
                         Rc * ((M_PI * (3.0 * rho_max + 1.0) + 9.5) * u + (double)kFastTrigErr);
    const double eps = 1.25 * (e_ref + e_apx) + 1e-6;
    const int one = 1 << fb;
    const int k = (int)std::ceil((double)one * eps + 0.5);
    if (k > one / 32) return fp;                 // the uncertain band would cover > 6 % of every cell: not worth it
    int kb = 1;
    while (kb < k) kb <<= 1;                     // power-of-two band: "within the band" becomes one AND
    fp.enabled = 1;
    fp.frac_bits = fb;
    fp.fmask = (one - 1) & ~(2 * kb - 1);
    const float magic_base = (float)(1.5 * (double)(1 << (23 - fb)));
    int mb_bits;
    std::memcpy(&mb_bits, &magic_base, 4);
    fp.mbk = mb_bits >> fb;
    fp.magic = magic_base + (float)kb / (float)one;
    fp.band = (float)kb / (float)one + 0.5f / (float)one;
    fp.t_dir = (float)(3.0 * (1.0 + eps) + 4.0 * u * Rc + 1e-4);
    fp.t_dir_neg = (float)(5.0 * (1.0 + eps) + 4.0 * u * Rc + 1e-4);
    fp.x2_min = (float)(3.0 * eps + 1e-3);       // the extended point's error is below 2 eps (twice the ray term)
    // reference: score 0 when trunc(e) <= -2 or >= W + 1 on either axis, i.e. e <= -2 or e >= W + 1
    fp.grid_w = h->grid.width; fp.grid_h = h->grid.height;
    fp.ghalf_x = (float)(0.5 * (h->grid.width + 3) + eps + 1e-3);
    fp.ghalf_y = (float)(0.5 * (h->grid.height + 3) + eps + 1e-3);
    fp.rho_lo = (float)h->ratio_lo; fp.rho_hi = (float)h->ratio_hi;
    fp.max_shift = (float)max_shift;
    fp.coord_hi = (float)(Cm - 1.0);
    fp.reach = (float)(Rc * (1.0 + 1e-6) + 4.0);
    fp.grid_min_dim = (float)std::min(h->grid.width, h->grid.height);
    fp.rho_abs = (float)(rho_max * (1.0 + 1e-6));
    fp.ang_room = 9.5f - h->max_abs_theta;      // kFastTrigErr is measured for |angle| <= 9.5
    plan_set_window(fp, x0, y0, w, hh, pitch);
    if (fp.half_x <= 0.0f || fp.half_y <= 0.0f) { fp.enabled = 0; return fp; }
    h->stats_eps = eps;
    return fp;
}

// ---- distance transform of the mirror (mcl_kernels.cuh: dt_*_kernel) ------------------------------------------------------
int run_distance_transform(mcl_engine* h, int thr)
{
    const int W = h->grid.width, H = h->grid.height;
    if (!h->dt_steps) CK(cudaMalloc((void**)&h->dt_steps, sizeof(uint16_t) * (size_t)W * H));
    dt_columns_kernel<<<(W + 127) / 128, 128, 0, h->stream>>>(h->map, W, H, h->grid.pitch, thr, h->dt_steps);
    CKL(h);
    dt_rows_kernel<<<(H + 127) / 128, 128, 0, h->stream>>>(W, H, h->dt_steps);
    CKL(h);
    return MCL_OK;
}

int refresh_likelihood_field(mcl_engine* h)
{
    if (!h->lf_dirty) return MCL_OK;
    int rc = run_distance_transform(h, 1);          // steps to the nearest OCCUPIED cell
    if (rc) return rc;
    lf_field_kernel<<<grid_for(h, (long long)h->grid.width * h->grid.height, 256), 256, 0, h->stream>>>(
        h->dt_steps, h->grid.width, h->grid.height, h->grid.pitch, h->map_lf);
    CKL(h);
    h->lf_dirty = false;
    return refresh_class_map(h, -kTabApron, -kTabApron, h->grid.width + 2 * kTabApron, h->grid.height + 2 * kTabApron);
}

// ---- table sensor path (mcl_table.cuh) ------------------------------------------------------------------------------
// Launch order on the engine's stream: bbox_kernel -> table_plan_kernel (one thread: window + budget, in device memory)
// -> score_table_kernel.  No host round trip: the host only needs to know whether the table pass is APPLICABLE, and
// takes that from the plan of the previous update (copied to pinned memory behind the kernel); the kernel itself
// verifies the plan and sends every evaluation down the exact path when it does not hold, so a stale hint costs time,
// never correctness.  The first scoring pass after the cloud was (re)initialised plans synchronously.
constexpr long long kTabMinParticles = 64;

// returns 1 (one window) or 2 (one window per batch) when the table kernel has been launched, 0 when the caller should use
// the other kernel families, < 0 on error
int run_score_table(mcl_engine* h, ScoreArgs& sa)
{
    const long long local = h->hi - h->lo;
    const int lanes = h->params.lanes_per_particle;
    const bool lf = h->params.sensor_mode == 1;
    const bool cand = (lf || (h->params.sensor_path == 0 && h->params.map_tile != 1 && (lanes == 0 || lanes == 1) &&
                              local >= kTabMinParticles)) && local < (1ll << 31) && h->num_beams > 0 && h->num_beams <= kTabMaxBeams && h->scan_finite &&
                      std::isfinite(h->max_range) && !std::getenv("MCL_NO_TABLE");
    if (!cand) return lf ? fail(h, MCL_ERR_INVALID, "the likelihood-field mode needs a finite scan of at most %d beams", kTabMaxBeams) : 0;
    if (lf) { int rc = refresh_likelihood_field(h); if (rc) return rc; }
    auto debug_plan = [&](const char* what, const TabPlan& d) {
        if (std::getenv("MCL_DEBUG_TABLE"))
            std::fprintf(stderr, "[mcl table] %s: variant %d (batch of %d) -> ok %d reason %d best %d misfits %d w %d h %d cap %d; built: entries %d "
                         "overflows %d windowless %d particles outside %d\n", what, d.variant, h->tab_batch, d.ok, d.reason, d.best, d.misfits,
                         d.w, d.h, d.cap_entries, h->tab_hint->build[0], h->tab_hint->build[1], h->tab_hint->build[3], h->tab_hint->build[4]);
    };
    if (h->tab_hint_pending && cudaEventQuery(h->ev_tab_hint) == cudaSuccess) {
        h->tab_hint_pending = false;
        const TabPlan& hp = h->tab_hint->plan;
        debug_plan("hint", hp);
        h->tab_ok = hp.ok != 0;
        const long long hint_batches = hp.batch ? (local + h->tab_batch - 1) / h->tab_batch : 1;
        if (hp.ok && h->tab_hint->build[1] * 50ll > (hp.batch ? hint_batches : 0)) {
            // the score table overflowed (the kernel scored those windows exactly): smaller batches, else another variant
            h->tab_ok = false;
            if (hp.batch && h->tab_batch > kTabBatchSmall) h->tab_batch >>= 1;
            else h->tab_excluded |= 1 << hp.variant;
        } else if (hp.best >= 0 && hp.best != hp.variant) {
            // follow what the last plan saw: one window when the cloud fits one, 16-bit classes when they fit, ...
            // (when this plan was not ok, tab_ok is false and the next pass plans synchronously)
            h->tab_variant = hp.best;
        }
        // a culling pass that finds nothing is tried less and less often (every 2nd ... 16th update); a hit resets that
        if (h->tab_hint->build[6]) h->cull_interval = h->tab_hint->build[5] > 0 ? 1 : std::min(16, h->cull_interval * 2);
    }
    if (h->tab_blocked > 0 && !lf) { --h->tab_blocked; return 0; }
    const size_t smem_total = (size_t)h->max_smem_optin - 1024;      // static shared memory of the kernel stays below 1 KB
    long long nbatches = 0;
    {
        const long long most = (local + kTabBatchSmall - 1) / kTabBatchSmall;
        if ((size_t)most > h->tab_bboxes_cap) {
            if (h->tab_bboxes) cudaFree(h->tab_bboxes);
            h->tab_bboxes = nullptr; h->tab_bboxes_cap = 0;
            CK(cudaMalloc((void**)&h->tab_bboxes, sizeof(int4) * (size_t)most));
            h->tab_bboxes_cap = (size_t)most;
        }
    }
    TabPlanIn in{};
    in.grid = h->grid;
    in.max_range = h->max_range; in.min_range = h->params.min_range; in.max_abs_theta = h->max_abs_theta;
    in.ratio_lo = h->ratio_lo; in.ratio_hi = h->ratio_hi;
    in.num_beams = h->num_beams;
    in.scan_finite = h->scan_finite ? 1 : 0;
    in.allow = 1;
    in.cull = (!lf && !std::getenv("MCL_NO_CULL")) ? 1 : 0;
    in.smem_total = (int)smem_total;
    in.smem_fixed = (int)table_fixed_smem(h->num_beams);
    const bool allow_batch = !lf && !std::getenv("MCL_NO_TABLE_BATCH");          // (the likelihood field keeps one window)
    const double rc6 = (double)h->max_range * (double)h->grid.cells_per_meter + 6.0;
    const long long k_budget = (long long)in.smem_total - in.smem_fixed - 64;       // for tab_batch_bytes
    auto plan = [&]() -> int {
        in.variant = h->tab_variant;
        in.excluded = h->tab_excluded;
        nbatches = (local + h->tab_batch - 1) / h->tab_batch;
        in.num_batches = allow_batch ? nbatches : 0;
        table_bbox_kernel<<<(int)nbatches, 256, 0, h->stream>>>(sa.x, sa.y, sa.th, sa.px, sa.py, sa.pth, h->lo, h->hi, h->tab_batch,
                                                                h->grid, rc6, k_budget, h->tab_box,
                                                                allow_batch ? h->tab_bboxes : nullptr);
        CKL(h);
        table_plan_kernel<<<1, 1, 0, h->stream>>>(in, h->tab_box, h->tab_plan, h->tab_build, h->deferred_counter,
                                                  h->count_gathers ? h->gather_counter : nullptr);
        CKL(h);
        return MCL_OK;
    };
    int rc = plan();
    if (rc) return rc;
    if (!h->tab_ok) {
        // no usable hint (first pass after an init / import, or the last plan did not allow the table): plan synchronously
        for (int attempt = 0; attempt < 6; ++attempt) {
            CK(cudaMemcpyAsync(&h->tab_hint->plan, h->tab_plan, sizeof(TabPlan), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->tab_hint_pending = false;
            const TabPlan& hp = h->tab_hint->plan;
            h->tab_ok = hp.ok != 0;
            debug_plan("plan", hp);
            if (h->tab_ok || hp.reason != 3) break;                            // (3: the window does not fit)
            if (hp.best >= 0) h->tab_variant = hp.best;                        // another variant is applicable
            else if (allow_batch && h->tab_batch > kTabBatchSmall) h->tab_batch >>= 1;     // 4096 -> 2048 -> 1024: smaller windows
            else break;
            rc = plan();
            if (rc) return rc;
        }
        if (!h->tab_ok && !lf) return 0;      // (likelihood-field mode: the kernel then scores every ray exactly, from the field)
        h->stats_eps = h->tab_hint->plan.eps;
    }
    const bool batch = h->tab_variant >= kTabBatch16, wide = (h->tab_variant & 1) != 0;
    TabArgs a{};
    a.x = sa.x; a.y = sa.y; a.th = sa.th; a.px = sa.px; a.py = sa.py; a.pth = sa.pth;
    a.score2 = sa.score2;
    a.lo = sa.lo; a.hi = sa.hi;
    a.beams = sa.beams; a.num_beams = sa.num_beams;
    a.grid = sa.grid;
    a.fast_cells = sa.fast_cells;
    a.lf_cells = lf ? h->map_lf : nullptr;
    a.cls = h->map_cls; a.pack = h->map_pack; a.cpitch = h->cpitch;
    a.plan = h->tab_plan;
    a.gather_counter = sa.gather_counter;
    a.deferred_counter = sa.deferred_counter;
    a.build_info = h->tab_build;
    a.bboxes = h->tab_bboxes;
    a.batch = h->tab_batch;
    a.cull = nullptr;
    if (!batch && in.cull && ++h->cull_wait >= h->cull_interval) {
        h->cull_wait = 0;
        a.cull = h->tab_cull;
    }
    if (a.cull) {
        table_cull_kernel<<<h->num_beams, 256, 0, h->stream>>>(h->tab_plan, sa.beams, h->num_beams, h->grid, h->map_cls, h->cpitch,
                                                               h->tab_cull, h->tab_build + 5);
        CKL(h);
    }
    const long long nunits = (local + 31) / 32;
    // one CTA per SM; small clouds get as many CTAs as their (unit, 32-beam word) pairs can keep busy
    const long long pairs = nunits * ((h->num_beams + 31) / 32);
    const int blocks = batch ? (int)std::min<long long>(nbatches, h->sm_count)
                             : (int)std::max<long long>(1, std::min<long long>((pairs + kTabWarps - 1) / kTabWarps, h->sm_count));
    if (!batch) {   // the units of the last partial round are split by beams and ADD to their particles' scores: zero those
        long long whole_units, items;
        int split;
        table_tail_split(nunits, (long long)blocks * kTabWarps, (h->num_beams + 31) / 32, whole_units, split, items);
        if (whole_units < nunits)
            CK(cudaMemsetAsync(a.score2 + a.lo + whole_units * 32, 0, sizeof(int32_t) * (size_t)(a.hi - a.lo - whole_units * 32),
                               h->stream));
    }
    auto launch = [&](auto kernel) -> int {
        { int rc2 = raise_smem(h, kernel, smem_total); if (rc2) return rc2; }
        kernel<<<blocks, kTabThreads, smem_total, h->stream>>>(a);
        CKL(h);
        return MCL_OK;
    };
    auto pick = [&](auto interp, auto count) -> int {
        constexpr bool I = decltype(interp)::value, C = decltype(count)::value;
        if (batch) return wide ? launch(score_table_kernel<I, C, true, true>) : launch(score_table_kernel<I, C, true, false>);
        return wide ? launch(score_table_kernel<I, C, false, true>) : launch(score_table_kernel<I, C, false, false>);
    };
    using T = std::true_type; using F = std::false_type;
    if (h->scan_interp) rc = h->count_gathers ? pick(T{}, T{}) : pick(T{}, F{});
    else rc = h->count_gathers ? pick(F{}, T{}) : pick(F{}, F{});
    if (rc) return rc;
    // the plan and the build summary follow the kernel to pinned memory: the hint for the next update
    CK(cudaMemcpyAsync(h->tab_hint, h->tab_hint_dev, sizeof(mcl_engine::TabHint), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev_tab_hint, h->stream));
    h->tab_hint_pending = true;
    return 1 + h->tab_variant;
}

int run_score(mcl_engine* h)
{
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map has not been called");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles: call mcl_init_* or mcl_import_particles");
    if (!h->have_scan) return fail(h, MCL_ERR_STATE, "no scan uploaded");
    const long long local = h->hi - h->lo;
    ScoreArgs a{};
    const PoseSoA& p = h->pose[h->cur];
    const PoseSoA& q = h->parent[h->cur];
    a.x = p.x; a.y = p.y; a.th = p.th; a.px = q.x; a.py = q.y; a.pth = q.th;
    a.score2 = h->score2;           // every rank scores, floors and normalises its own slice: scores never travel
    a.fast_cells = h->map_fast;
    a.lo = h->lo; a.hi = h->hi;
    a.beams = h->beams; a.num_beams = h->num_beams;
    a.grid = h->grid;
    a.gather_counter = h->gather_counter;
    a.deferred_counter = h->deferred_counter;
    if (local > 0) {
        const int tr = run_score_table(h, a);       // (its plan kernel zeroes the counters)
        if (tr < 0) return tr;
        if (tr >= 1) {
            int rc = join_pushes(h);
            if (rc) return rc;
            h->stats.lanes_per_particle = 1;
            h->stats.map_tile_used = tr >= 3 ? 5 : 4;
            h->stats.table_variant = tr - 1;
            h->stats.sensor_path = 3;
            h->stats.fast_eps = h->stats_eps;
            h->stats.evals = local * (long long)h->num_beams;
            h->have_scores = true;
            return MCL_OK;
        }
    }
    if (h->count_gathers) CK(cudaMemsetAsync(h->gather_counter, 0, sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->deferred_counter, 0, sizeof(unsigned long long), h->stream));

    // lanes per particle: enough particle groups to fill the machine (148 SMs x 8 CTAs x (256/G) slots)
    int G = h->params.lanes_per_particle;
    if (G != 1 && G != 2 && G != 4 && G != 8 && G != 16 && G != 32) {
        G = 32;
        while (G > 1 && local * G / 2 >= (long long)h->sm_count * 2048 * 2) G >>= 1;
    }

    // shared-memory map tile: window = bounding box of the cloud (poses and parents) +- (max range + 2 cells)
    size_t smem = (size_t)h->num_beams * sizeof(Beam);          // == sizeof(FastBeam) per beam
    bool tile = false, batch = false;
    size_t batch_tile_bytes = 0;
    long long batch_max_dim = 0;
    if (h->params.map_tile != 1 && local > 0 && h->num_beams > 0) {
        int* box = h->bbox;
        const int init_box[4] = {0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000};
        std::memcpy(h->host_bbox, init_box, sizeof(init_box));
        CK(cudaMemcpyAsync(box, h->host_bbox, sizeof(init_box), cudaMemcpyHostToDevice, h->stream));
        bbox_kernel<<<grid_for(h, local, 256), 256, 0, h->stream>>>(p.x, p.y, q.x, q.y, h->lo, h->hi, box);
        CKL(h);
        CK(cudaMemcpyAsync(h->host_bbox, box, sizeof(init_box), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        auto unorder = [](int i) { int b = i >= 0 ? i : i ^ 0x7fffffff; float f; std::memcpy(&f, &b, 4); return f; };
        const float mnx = unorder(h->host_bbox[0]), mny = unorder(h->host_bbox[1]);
        const float mxx = unorder(h->host_bbox[2]), mxy = unorder(h->host_bbox[3]);
        if (std::isfinite(mnx) && std::isfinite(mny) && std::isfinite(mxx) && std::isfinite(mxy)) {
            const double cpm = h->grid.cells_per_meter;
            const double reach = (double)h->max_range * cpm + 3.0;
            const double cx0 = std::floor(((double)mnx - h->grid.origin_x) * cpm - reach);
            const double cy0 = std::floor(((double)mny - h->grid.origin_y) * cpm - reach);
            const double cx1 = std::ceil(((double)mxx - h->grid.origin_x) * cpm + reach);
            const double cy1 = std::ceil(((double)mxy - h->grid.origin_y) * cpm + reach);
            // clip to the grid plus a 2-cell zero margin: endpoints on the map's outer wall stay interior to the
            // window (cells outside the grid are staged as 0, OccupancyGrid::logOdds' out-of-grid value)
            long long x0 = (long long)std::max(cx0, -2.0), y0 = (long long)std::max(cy0, -2.0);
            long long x1 = (long long)std::min(cx1, (double)h->grid.width + 1);
            long long y1 = (long long)std::min(cy1, (double)h->grid.height + 1);
            if (x1 >= x0 && y1 >= y0) {
                x0 = (x0 >= 0) ? (x0 & ~3ll) : -(((-x0) + 3) & ~3ll);   // 4-byte aligned, rounding down
                const long long tw = x1 - x0 + 1, th = y1 - y0 + 1;
                long long pitch = (tw + 3) & ~3ll;
                if (((pitch >> 2) & 1) == 0) pitch += 4;     // odd number of 4-byte words per row: spreads rows over banks
                const size_t bytes = (size_t)pitch * th;
                // the exact pass of the two-pass path also holds its compaction queues in shared memory
                const size_t extra2 = h->params.sensor_path != 1
                                          ? deferred_smem_bytes(h->num_beams, MCL_DEF_THREADS / 32) - smem : 0;
                const size_t budget = (size_t)h->max_smem_optin - smem - extra2 - 1024;
                if (bytes <= budget / (h->params.map_tile == 2 ? 1 : 2) || (h->params.map_tile == 2 && bytes <= budget)) {
                    tile = true;
                    a.tile_x0 = (int)x0; a.tile_y0 = (int)y0; a.tile_w = (int)tw; a.tile_h = (int)th;
                    a.tile_pitch = (int)pitch;
                    smem += bytes;
                }
            }
        }
        // The cloud as a whole does not fit one tile: try one window per batch of consecutive particles (they are
        // spatial neighbours after mcl_init_uniform, and systematic resampling preserves the order).
        if (!tile && G == 1 && h->params.map_tile != 2 && std::isfinite(h->max_range)) {
            const long long nb = (local + kBatchParticles - 1) / kBatchParticles;
            if ((size_t)nb > h->windows_cap) {
                if (h->windows) cudaFree(h->windows);
                h->windows = nullptr; h->windows_cap = 0;
                CK(cudaMalloc((void**)&h->windows, sizeof(BatchWindow) * (size_t)nb));
                h->windows_cap = (size_t)nb;
            }
            const size_t base2 = std::max(smem, deferred_smem_bytes(h->num_beams, kBatchDefThreads / 32));
            const long long budget = (long long)h->max_smem_optin - (long long)base2 - 1024;
            if (budget > 4096) {
                h->host_bbox[0] = 0; h->host_bbox[1] = 0; h->host_bbox[2] = 0;
                CK(cudaMemcpyAsync(h->bbox, h->host_bbox, 12, cudaMemcpyHostToDevice, h->stream));
                batch_window_kernel<<<(int)nb, 256, 0, h->stream>>>(p.x, p.y, q.x, q.y, h->lo, h->hi, h->grid,
                                                                   (double)h->max_range * h->grid.cells_per_meter + 3.0,
                                                                   (int)budget, h->windows, h->bbox);
                CKL(h);
                CK(cudaMemcpyAsync(h->host_bbox, h->bbox, 12, cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                const int max_bytes = h->host_bbox[0], misfits = h->host_bbox[1];
                batch_max_dim = h->host_bbox[2];
                if (max_bytes > 0 && (long long)misfits * 50 <= nb) {     // at most 2 % of the batches without a tile
                    batch = true;
                    tile = true;
                    a.windows = h->windows;
                    a.num_batches = nb;
                    batch_tile_bytes = (size_t)max_bytes;
                    a.tile_x0 = -4; a.tile_y0 = -4; a.tile_w = h->grid.width + 8; a.tile_h = h->grid.height + 8;   // eps only
                    a.tile_pitch = 4;
                }
            }
        }
        if (h->params.map_tile == 2 && !tile)
            return fail(h, MCL_ERR_INVALID, "map_tile=2 forced but the cloud's window does not fit in shared memory");
    }
    if (h->params.sensor_path != 1 && local > 0)
        a.fast = batch ? fast_plan(h, a.tile_x0, a.tile_y0, a.tile_w, a.tile_h, a.tile_pitch, batch_max_dim)
               : tile  ? fast_plan(h, a.tile_x0, a.tile_y0, a.tile_w, a.tile_h, a.tile_pitch, std::max(a.tile_w, a.tile_h))
                       : fast_plan(h, 0, 0, h->grid.width, h->grid.height, h->grid.pitch,
                                   std::max(h->grid.width, h->grid.height));
    const bool fast = a.fast.enabled != 0;
    int rc;
    if (fast) {
        // two passes: certified float evaluation, then the literal restatement for the beams it deferred
        const int iters = (h->num_beams + G - 1) / G;
        const size_t need = (size_t)((iters + 31) / 32) * (size_t)(local * G) * sizeof(uint32_t);
        if (need > h->masks_bytes) {
            if (h->masks) cudaFree(h->masks);
            h->masks = nullptr; h->masks_bytes = 0;
            CK(cudaMalloc((void**)&h->masks, need));
            h->masks_bytes = need;
        }
        a.masks = h->masks;
        rc = launch_score(h, G, tile, batch, a, 1, smem + batch_tile_bytes, local);
        if (rc) return rc;
        const size_t smem2 = smem - (size_t)h->num_beams * sizeof(Beam) +
                             deferred_smem_bytes(h->num_beams, (batch ? kBatchDefThreads : MCL_DEF_THREADS) / 32);
        rc = launch_score(h, G, tile, batch, a, 2, smem2 + batch_tile_bytes, local);
    } else {
        rc = launch_score(h, G, tile, batch, a, 0, smem + batch_tile_bytes, local);
    }
    if (rc) return rc;
    // (the stream now waits for this rank's pose pushes; the normaliser's first barrier then makes every rank's pushes
    // complete before anyone starts the next update)
    rc = join_pushes(h);
    if (rc) return rc;
    h->stats.lanes_per_particle = G;
    h->stats.map_tile_used = batch ? 3 : (tile ? 2 : 1);
    h->stats.sensor_path = fast ? 2 : 1;
    h->stats.fast_eps = fast ? h->stats_eps : 0.0;
    h->stats.evals = local * (long long)h->num_beams;
    h->have_scores = true;
    return MCL_OK;
}

int run_normalize(mcl_engine* h)
{
    if (!h->have_scores) return fail(h, MCL_ERR_STATE, "mcl_score has not run");
    double* w = h->weight[h->wcur];
    const long long lo = h->lo, hi = h->hi, local = hi - lo;
    const int g = grid_for(h, local, 256);
    if (h->params.weight_mode == 1) {
        // extension: w = exp(beta (s - max s)) / sum  (max / log-sum-exp normalisation)
        if (h->world > 1) return fail(h, MCL_ERR_STATE, "weight_mode 1 runs on single-GPU engines");
        CK(cudaMemsetAsync(h->bbox, 0x80, sizeof(int), h->stream));          // 0x80808080: below every score
        score_max_kernel<<<g, 256, 0, h->stream>>>(h->score2, h->n, h->bbox);
        CKL(h);
        lse_kernel<<<g, 256, 0, h->stream>>>(h->score2, w, h->n, h->bbox, 0.5 * h->params.lse_beta);
        CKL(h);
    } else if (local > 0) {
        floor_kernel<<<g, 256, 0, h->stream>>>(h->score2 + lo, w + lo, local, h->params.weight_floor);
        CKL(h);
    }
    prof_mark(h, "normalise:floor");
    int rc = seq_total(h, h->wcur, h->ess_acc);          // (the walk also zeroes the sum of squares the divide accumulates)
    if (rc) return rc;
    if (local > 0) {
        xdivide_kernel<<<g, 256, 0, h->stream>>>(w, lo, hi, h->total, h->ess_acc);
        CKL(h);
    }
    return MCL_OK;
}

int run_estimate(mcl_engine* h)
{
    const PoseSoA& p = h->pose[h->cur];
    const long long chunk_lo = h->lo / kEstChunk, chunk_hi = (h->hi + kEstChunk - 1) / kEstChunk;
    if (chunk_hi > chunk_lo) {
        xestimate_partial_kernel<<<(int)(chunk_hi - chunk_lo), kEstBlock, 0, h->stream>>>(
            p.x, p.y, p.th, h->weight[h->wcur], h->n, chunk_lo, h->xp, h->xl.est, h->xl.est_w2);
        CKL(h);
    }
    int rc = xbarrier(h, 2);
    if (rc) return rc;
    xestimate_final_kernel<<<1, kEstBlock, 0, h->stream>>>((const double4*)(h->xblock + h->xl.est),
                                                          (const double*)(h->xblock + h->xl.est_w2), h->est_count,
                                                          h->est_out, h->ess_acc);
    CKL(h);
    return MCL_OK;
}

int run_action(mcl_engine* h, const mcl_action_t* act, int64_t utime, const float* noise_dev, bool from_index)
{
    ActionArgs a{};
    const PoseSoA& src = h->pose[h->cur];
    const int dst_i = h->cur ^ 1;
    a.sx = src.x; a.sy = src.y; a.sth = src.th;
    a.src_index = from_index ? h->idx : nullptr;
    a.dx = h->pose[dst_i].x; a.dy = h->pose[dst_i].y; a.dth = h->pose[dst_i].th;
    a.dpx = h->parent[dst_i].x; a.dpy = h->parent[dst_i].y; a.dpth = h->parent[dst_i].th;
    a.lo = h->lo; a.hi = h->hi;
    a.rot1 = act->rot1; a.trans = act->trans; a.rot2 = act->rot2;
    a.s1 = act->rot1_std; a.st = act->trans_std; a.s2 = act->rot2_std;
    a.moved = act->moved;
    a.noise = noise_dev;
    a.seed = h->seed;
    a.update_no = h->update_no;
    action_kernel<<<grid_for(h, h->hi - h->lo, 256), 256, 0, h->stream>>>(a);
    CKL(h);
    // every rank needs the complete new cloud: the next resampling gathers parents from anywhere in it
    if (h->world > 1 && h->peer_push) {
        // copy engines push this rank's slice of (x, y, theta) into every peer's buffer while the SMs score;
        // the score exchange that follows waits for this rank's pushes, which makes it the cross-rank barrier
        CK(cudaEventRecord(h->ev_action_done, h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev_action_done, 0));
        const size_t off = (size_t)(3 * dst_i) * h->n + (size_t)h->lo;
        const size_t width = (size_t)(h->hi - h->lo) * sizeof(float), pitch = (size_t)h->n * sizeof(float);
        for (int r = 0; r < h->world; ++r) {
            if (r == h->rank) continue;
            CK(cudaMemcpy2DAsync(h->peer_pose_block[r] + off, pitch, h->pose_block + off, pitch, width, 3,
                                 cudaMemcpyDeviceToDevice, h->copy_stream));
        }
        CK(cudaEventRecord(h->ev_push_done, h->copy_stream));
        h->push_pending = true;
        ++h->collectives;
    } else {
        for (float* arr : {h->pose[dst_i].x, h->pose[dst_i].y, h->pose[dst_i].th}) {
            int rc = exchange_slices(h, arr, sizeof(float));
            if (rc) return rc;
        }
    }
    h->cur = dst_i;
    // action_model.cpp:92-93: parent keeps the old pose (and its utime), pose.utime = ActionModel::utime_
    h->parent_utime = h->pose_utime;
    h->pose_utime = h->params.legacy_equal_utime ? h->pose_utime : utime;
    return MCL_OK;
}

int upload_noise(mcl_engine* h, const float* noise3n, const float** dev_out)
{
    *dev_out = nullptr;
    if (!noise3n) return MCL_OK;
    if (!h->noise) CK(cudaMalloc((void**)&h->noise, sizeof(float) * 3 * (size_t)h->n));
    CK(cudaMemcpyAsync(h->noise, noise3n, sizeof(float) * 3 * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    *dev_out = h->noise;
    return MCL_OK;
}

// children / clo / chi: the draws U_m = r + m / children, m in [clo, chi), this call resolves (the filter's resampling:
// children = N and the rank's slice; a weighted export: any number of draws); indices go to idx_out[m].
int run_resample_indices(mcl_engine* h, double r, int wbuf, long long children = -1, long long clo = 0, long long chi = 0,
                         int32_t* idx_out = nullptr)
{
    int rc = seq_total(h, wbuf);
    if (rc) return rc;
    // materialise the exact running sum only where this rank's children draw from, then search
    const long long n = h->n, n1 = h->n1, n2 = h->n2;
    if (children < 0) { children = n; clo = h->lo; chi = h->hi; idx_out = h->idx; }
    const long long local = chi - clo;
    xresample_range_kernel<<<1, 32, 0, h->stream>>>(h->cin2, n2, children, r, clo, chi, h->xrange, h->overruns);
    CKL(h);
    const long long groups_cap = (h->world == 1 || children != n) ? n2
                               : std::min<long long>(n2, 2 * ((local + kSliceAlign - 1) / kSliceAlign) + 64);
    xseq_group_expand_kernel<<<(int)((n2 + 127) / 128), 128, 0, h->stream>>>(h->q0, h->q1, n1, h->cin2, h->opened, h->cin1,
                                                                         h->xrange);
    CKL(h);
    const long long tiles_cap = std::max<long long>(1, groups_cap * kL2 / kSeqTileChunks);
    xseq_materialize_kernel<<<(int)std::min<long long>(tiles_cap, (long long)h->sm_count * 32), 128, 0, h->stream>>>(
        n, n1, h->cin1, h->cum, h->xrange, h->xp, h->xl.w[wbuf]);
    CKL(h);
    if (local > 0) {
        xresample_search_kernel<<<grid_for(h, (local + kSearchRun - 1) / kSearchRun, 128), 128, 0, h->stream>>>(
            h->cum, n, children, r, clo, chi, h->xrange, idx_out, h->overruns);
        CKL(h);
    }
    return MCL_OK;
}

// What a blocking call reads back, gathered on the device so that it is ONE copy (each tiny D2H copy costs the host
// several microseconds): counters[0..5] = overruns, gathers, fall-back chunks, (double) weight sum, (double) sum w^2,
// deferred evaluations; counters[6] = barrier time-out flag; est = the pose estimate.
__global__ void readback_kernel(const unsigned long long* overruns, const unsigned long long* gathers,
                                const long long* fallbacks, const double* total, const double* ess,
                                const unsigned long long* deferred, const int* err, const float* est, const int* tab_build,
                                Readback* out)
{
    Readback r;
    r.counters[0] = *overruns; r.counters[1] = *gathers; r.counters[2] = (unsigned long long)*fallbacks;
    r.counters[3] = (unsigned long long)__double_as_longlong(*total);
    r.counters[4] = (unsigned long long)__double_as_longlong(*ess);
    r.counters[5] = *deferred; r.counters[6] = (unsigned long long)(unsigned)*err;
    for (int i = 0; i < 4; ++i) r.est[i] = est[i];
    r.culled = tab_build[5]; r.pad = 0;
    *out = r;
}

int read_back(mcl_engine* h, bool counters, bool estimate, int64_t utime)
{
    readback_kernel<<<1, 1, 0, h->stream>>>(h->overruns, h->gather_counter, (const long long*)h->fallbacks, h->total, h->ess_acc,
                                            h->deferred_counter, (const int*)(h->xblock + h->xl.err), h->est_out, h->tab_build,
                                            h->readback_dev);
    CKL(h);
    CK(cudaMemcpyAsync(h->readback_host, h->readback_dev, sizeof(Readback), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const Readback& r = *h->readback_host;
    if ((int)(r.counters[6] & 0xffffffffu) != 0)
        return fail(h, MCL_ERR_COMM, "a cross-rank barrier timed out (a peer rank stopped)");
    if (counters) {
        h->stats.deferred_evals = (int64_t)r.counters[5];
        h->stats.resample_overruns = (int64_t)r.counters[0];
        h->stats.gathers = h->count_gathers ? (int64_t)r.counters[1] : -1;
        h->stats.seq_fallback_chunks = (int64_t)r.counters[2];
        h->stats.culled_beams = h->stats.sensor_path == 3 ? r.culled : 0;
        double d;
        std::memcpy(&d, &r.counters[3], 8);
        h->stats.weight_sum = d;
        std::memcpy(&d, &r.counters[4], 8);
        h->stats.effective_sample_size = d > 0 ? 1.0 / d : 0.0;
    }
    if (estimate) {
        h->last_estimate.x = r.est[0];
        h->last_estimate.y = r.est[1];
        h->last_estimate.theta = r.est[2];
        h->last_estimate.utime = utime;
    }
    return MCL_OK;
}

int read_counters(mcl_engine* h) { return read_back(h, true, false, 0); }
int fetch_estimate(mcl_engine* h, int64_t utime) { return read_back(h, false, true, utime); }

// The five stages of ParticleFilter::updateFilter on the engine's stream, no host synchronisation unless the tile
// heuristic needs the cloud's bounding box.
// scan (optional): host scan buffers to prepare and upload between the action step and the sensor stage, so that the
// host's share of it (compaction, interpolation ratios) runs while the GPU resamples and moves the cloud.
struct HostScan { const float* ranges; const float* thetas; const int64_t* times; int nb; long long t_begin, t_end; };
int enqueue_update(mcl_engine* h, const mcl_action_t* a, int64_t utime, double r, const float* noise_dev,
                   const HostScan* scan = nullptr)
{
    h->launches = 0;
    h->collectives = 0;
    cudaEventRecord(h->ev[0], h->stream);
    prof_mark(h, "update:begin");
    int rc = run_resample_indices(h, r, h->wcur);
    if (rc) return rc;
    prof_mark(h, "resample:range+expand+materialise+search");
    cudaEventRecord(h->ev[1], h->stream);
    rc = run_action(h, a, utime, noise_dev, true);
    if (rc) return rc;
    prof_mark(h, "action");
    if (scan) {
        rc = prepare_scan(h, scan->ranges, scan->thetas, scan->times, scan->nb, scan->t_begin, scan->t_end);
        if (rc) return rc;
    }
    cudaEventRecord(h->ev[2], h->stream);
    rc = run_score(h);
    if (rc) return rc;
    prof_mark(h, "score (+ join pose pushes)");
    cudaEventRecord(h->ev[3], h->stream);
    rc = run_normalize(h);
    if (rc) return rc;
    prof_mark(h, "normalise:divide");
    cudaEventRecord(h->ev[4], h->stream);
    rc = run_estimate(h);
    if (rc) return rc;
    prof_mark(h, "estimate (partials, barrier, final)");
    cudaEventRecord(h->ev[5], h->stream);
    h->stats.kernel_launches = h->launches;
    h->stats.collectives = h->collectives;
    ++h->update_no;
    ++h->stats.updates;
    return MCL_OK;
}

void free_all(mcl_engine* h)
{
    auto F = [](void* p) { if (p) cudaFree(p); };
    for (int b = 0; b < 2; ++b) { F(h->parent[b].x); F(h->parent[b].y); F(h->parent[b].th); }
    for (size_t r = 0; r < h->peer_pose_block.size(); ++r)
        if (h->peer_pose_block[r] && h->peer_pose_block[r] != h->pose_block) cudaIpcCloseMemHandle(h->peer_pose_block[r]);
    for (int r = 0; r < kMaxPeers; ++r)
        if (h->xp.base[r] && h->xp.base[r] != h->xblock) cudaIpcCloseMemHandle(h->xp.base[r]);
    F(h->pose_block); F(h->barrier_word); F(h->xblock); F(h->fb_count); F(h->xrange);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_action_done) cudaEventDestroy(h->ev_action_done);
    if (h->ev_push_done) cudaEventDestroy(h->ev_push_done);
    F(h->tile_sums); F(h->tile_excl);
    F(h->score_block); F(h->idx); F(h->cum); F(h->sums); F(h->cin1); F(h->cin2); F(h->total);
    F(h->opened); F(h->fallbacks); F(h->overruns); F(h->gather_counter); F(h->deferred_counter); F(h->masks); F(h->windows); F(h->map_beams); F(h->map_counts); F(h->map_flag);
    if (h->map_beams_host) cudaFreeHost(h->map_beams_host);
    if (h->map_flag_host) cudaFreeHost(h->map_flag_host);
    F(h->ess_acc); F(h->bbox); F(h->est_out); F(h->map); F(h->map_fast); F(h->map_lf); F(h->map_cls); F(h->map_pack); F(h->dt_steps); F(h->beams); F(h->noise); F(h->staging);
    if (h->est_host) cudaFreeHost(h->est_host);
    if (h->beams_host) cudaFreeHost(h->beams_host);
    if (h->host_bbox) cudaFreeHost(h->host_bbox);
    if (h->host_bbox_init) cudaFreeHost(h->host_bbox_init);
    if (h->tab_hint) cudaFreeHost(h->tab_hint);
    if (h->ev_tab_hint) cudaEventDestroy(h->ev_tab_hint);
    F(h->tab_hint_dev); F(h->tab_box); F(h->tab_bboxes); F(h->tab_cull);
    if (h->readback_host) cudaFreeHost(h->readback_host);
    if (h->readback_dev) cudaFree(h->readback_dev);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
}

}  // namespace

// =====================================================================================================================
extern "C" {

void mcl_default_params(mcl_params* p)
{
    std::memset(p, 0, sizeof(*p));
    p->min_range = 0.15f;
    p->weight_floor = 0.001;
    p->init_std = 0.01;
    p->legacy_equal_utime = 0;
    p->lanes_per_particle = 0;
    p->map_tile = 0;
    p->sensor_path = 0;
    p->weight_mode = 0;
    p->sensor_mode = 0;
    p->lse_beta = 0.05;
}

const char* mcl_last_error(const mcl_engine* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int mcl_create(const mcl_params* params, int64_t num_particles, int device, mcl_engine** out)
{
    mcl_engine* h = nullptr;
    if (!out) return fail(h, MCL_ERR_INVALID, "out is null");
    *out = nullptr;
    if (num_particles < 2 || num_particles > 0x7fffffffLL)     // particle_filter.cpp:11 asserts > 1; indices are int32
        return fail(h, MCL_ERR_INVALID, "num_particles must be in [2, 2^31)");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return fail(h, MCL_ERR_NO_DEVICE, "no CUDA device %d (found %d): this engine has no CPU fallback", device, count);
    }
    h = new (std::nothrow) mcl_engine();
    if (!h) return fail(nullptr, MCL_ERR_INVALID, "out of host memory");
    if (params) h->params = *params; else mcl_default_params(&h->params);
    h->device = device;
    h->prof = std::getenv("MCL_PROFILE") != nullptr;
    h->n = num_particles;
    h->lo = 0; h->hi = num_particles;
    auto bail = [&](int rc) { g_last_error = h->err; free_all(h); delete h; return rc; };
#define CKB(call)                                                                                                    \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess) {                                                                                    \
            fail(h, MCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));                                  \
            return bail(MCL_ERR_CUDA);                                                                               \
        }                                                                                                            \
    } while (0)
    CKB(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKB(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        fail(h, MCL_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
             prop.minor);
        return bail(MCL_ERR_NO_DEVICE);
    }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    CKB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) CKB(cudaEventCreate(&e));
    const size_t n = (size_t)h->n;
    h->n1 = (h->n + kL1 - 1) / kL1;
    h->n2 = (h->n1 + kL2 - 1) / kL2;
    for (int b = 0; b < 2; ++b) {
        if (b == 0) CKB(cudaMalloc((void**)&h->pose_block, 4 * n * 6));
        h->pose[b].x = h->pose_block + (size_t)(3 * b + 0) * n;
        h->pose[b].y = h->pose_block + (size_t)(3 * b + 1) * n;
        h->pose[b].th = h->pose_block + (size_t)(3 * b + 2) * n;
        CKB(cudaMalloc((void**)&h->parent[b].x, 4 * n)); CKB(cudaMalloc((void**)&h->parent[b].y, 4 * n));
        CKB(cudaMalloc((void**)&h->parent[b].th, 4 * n));
    }
    CKB(cudaMalloc((void**)&h->score_block, 4 * n));
    h->score2 = h->score_block;
    CKB(cudaMalloc((void**)&h->idx, 4 * n));
    CKB(cudaMalloc((void**)&h->cum, 8 * n));
    const size_t n1 = (size_t)h->n1, n2 = (size_t)h->n2;
    CKB(cudaMalloc((void**)&h->sums, 8 * n1)); CKB(cudaMalloc((void**)&h->cin1, 8 * n1));
    CKB(cudaMalloc((void**)&h->tile_sums, 8 * (n1 / kSeqTileChunks + 1)));
    CKB(cudaMalloc((void**)&h->tile_excl, 8 * (n1 / kSeqTileChunks + 1)));
    CKB(cudaMalloc((void**)&h->cin2, 8 * n2));
    CKB(cudaMalloc((void**)&h->opened, 4 * n2));
    h->est_count = (int)((h->n + kEstChunk - 1) / kEstChunk);
    {   // the exchange block: everything another rank may write or read (mcl_shard.cuh)
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        XLayout& L = h->xl;
        size_t off = 0;
        L.w[0] = off; off = al(off + 8 * n);
        L.w[1] = off; off = al(off + 8 * n);
        L.q0 = off; off = al(off + 8 * n1);
        L.q1 = off; off = al(off + 8 * n1);
        L.eb = off; off = al(off + 4 * n1);
        L.g0 = off; off = al(off + 8 * n2);
        L.g1 = off; off = al(off + 8 * n2);
        L.ge = off; off = al(off + 4 * n2);
        L.tot = off; off = al(off + 8 * kMaxPeers);
        L.est = off; off = al(off + sizeof(double4) * (size_t)h->est_count);
        L.est_w2 = off; off = al(off + 8 * (size_t)h->est_count);
        L.fbraw = off; off = al(off + 8 * (size_t)kMaxPeers * kFbSlots * kL1);
        L.flags = off; off = al(off + 4 * (size_t)kXFlagSlots * kMaxPeers);
        L.err = off; off = al(off + 4);
        L.bytes = off;
        CKB(cudaMalloc((void**)&h->xblock, L.bytes));
        CKB(cudaMemset(h->xblock + L.q0, 0, L.bytes - L.q0));
        h->weight[0] = (double*)(h->xblock + L.w[0]); h->weight[1] = (double*)(h->xblock + L.w[1]);
        h->q0 = (long long*)(h->xblock + L.q0); h->q1 = (long long*)(h->xblock + L.q1); h->ebias = (int*)(h->xblock + L.eb);
        h->g0 = (long long*)(h->xblock + L.g0); h->g1 = (long long*)(h->xblock + L.g1); h->gebias = (int*)(h->xblock + L.ge);
        h->xp.world = 1; h->xp.rank = 0;
        for (int r = 0; r < kMaxPeers; ++r) h->xp.base[r] = nullptr;
        h->xp.base[0] = h->xblock;
        h->xp.lo[0] = 0;
        for (int r = 1; r <= kMaxPeers; ++r) h->xp.lo[r] = h->n;
    }
    CKB(cudaMalloc((void**)&h->fb_count, 4)); CKB(cudaMemset(h->fb_count, 0, 4));
    CKB(cudaMalloc((void**)&h->xrange, 16)); CKB(cudaMemset(h->xrange, 0, 16));
    CKB(cudaMalloc((void**)&h->total, 8)); CKB(cudaMalloc((void**)&h->fallbacks, 8));
    CKB(cudaMalloc((void**)&h->overruns, 8)); CKB(cudaMalloc((void**)&h->gather_counter, 8));
    CKB(cudaMalloc((void**)&h->deferred_counter, 8)); CKB(cudaMemset(h->deferred_counter, 0, 8));
    CKB(cudaMalloc((void**)&h->ess_acc, 8)); CKB(cudaMalloc((void**)&h->bbox, 16));
    CKB(cudaMemset(h->total, 0, 8)); CKB(cudaMemset(h->fallbacks, 0, 8)); CKB(cudaMemset(h->overruns, 0, 8));
    CKB(cudaMemset(h->gather_counter, 0, 8)); CKB(cudaMemset(h->ess_acc, 0, 8));
    CKB(cudaMalloc((void**)&h->est_out, 16));
    CKB(cudaMallocHost((void**)&h->est_host, 16));
    CKB(cudaMallocHost((void**)&h->host_bbox, 16));
    CKB(cudaMallocHost((void**)&h->host_bbox_init, 16));
    h->host_bbox_init[0] = h->host_bbox_init[1] = 0x7fffffff;
    h->host_bbox_init[2] = h->host_bbox_init[3] = (int)0x80000000;
    CKB(cudaMalloc((void**)&h->tab_box, 96));
    CKB(cudaMemset(h->tab_box, 0, 96));
    {
        const int arm[2] = {0x7fffffff, (int)0x80000000};          // heading range (table_bbox_kernel)
        CKB(cudaMemcpy(h->tab_box + 16, arm, 8, cudaMemcpyHostToDevice));
    }
    CKB(cudaMalloc((void**)&h->tab_cull, kTabMaxBeams + 1));
    CKB(cudaMemset(h->tab_cull, 0, kTabMaxBeams + 1));
    CKB(cudaMemcpy(h->tab_box, h->host_bbox_init, 16, cudaMemcpyHostToDevice));
    // the plan and the build summary share one buffer (same layout as the pinned TabHint): one copy brings both back
    CKB(cudaMalloc((void**)&h->tab_hint_dev, sizeof(mcl_engine::TabHint)));
    CKB(cudaMemset(h->tab_hint_dev, 0, sizeof(mcl_engine::TabHint)));
    h->tab_plan = &h->tab_hint_dev->plan;
    h->tab_build = h->tab_hint_dev->build;
    CKB(cudaMallocHost((void**)&h->tab_hint, sizeof(mcl_engine::TabHint)));
    std::memset(h->tab_hint, 0, sizeof(mcl_engine::TabHint));
    CKB(cudaEventCreateWithFlags(&h->ev_tab_hint, cudaEventDisableTiming));
    CKB(cudaMallocHost((void**)&h->readback_host, 128));
    CKB(cudaMalloc((void**)&h->readback_dev, 128));
#undef CKB
    h->stats.num_particles = h->n;
    h->stats.local_particles = h->n;
    h->stats.gathers = -1;
    *out = h;
    return MCL_OK;
}

void mcl_destroy(mcl_engine* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
#ifdef MCL_FAST_DIAG
    {
        unsigned long long d[8];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(d, g_fast_diag, sizeof(d));
        if (d[0])
            fprintf(stderr, "fast-pass diag: evals %llu  frac %.4f  outside-window %.4f  dir-band %.4f  x2-negative %.4f  deferred %.4f\n",
                    d[0], (double)d[1] / d[0], (double)d[2] / d[0], (double)d[3] / d[0], (double)d[4] / d[0], (double)d[5] / d[0]);
    }
#endif
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->prof && h->prof_updates > 0 && h->rank == 0) {
        fprintf(stderr, "MCL_PROFILE rank 0 of %d, %lld particles, %ld blocking updates, mean ms per part:\n", h->world,
                (long long)h->n, h->prof_updates);
        for (size_t i = 1; i < h->prof_sum.size(); ++i)
            fprintf(stderr, "  %-48s %8.4f\n", h->prof_label[i], h->prof_sum[i] / (double)h->prof_updates);
    }
    for (auto e : h->prof_ev) cudaEventDestroy(e);
#ifdef MCL_WITH_NCCL
    if (h->comm) ncclCommDestroy(h->comm);
#endif
    free_all(h);
    delete h;
}

void* mcl_stream(mcl_engine* h) { return h ? (void*)h->stream : nullptr; }

int mcl_sync(mcl_engine* h)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

int mcl_comm_unique_id(void* id128_out)
{
    mcl_engine* h = nullptr;
#ifdef MCL_WITH_NCCL
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return fail(h, MCL_ERR_COMM, "ncclGetUniqueId failed");
    std::memcpy(id128_out, &id, 128);
    return MCL_OK;
#else
    (void)id128_out;
    return fail(h, MCL_ERR_COMM, "library built without NCCL");
#endif
}

int mcl_comm_init(mcl_engine* h, const void* id128, int rank, int world)
{
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(h, MCL_ERR_INVALID, "bad comm arguments");
    if (world > kMaxPeers) return fail(h, MCL_ERR_INVALID, "at most %d ranks (one NVLink domain)", kMaxPeers);
#ifdef MCL_WITH_NCCL
    CK(cudaSetDevice(h->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    if (ncclCommInitRank(&h->comm, world, id, rank) != ncclSuccess) return fail(h, MCL_ERR_COMM, "ncclCommInitRank failed");
    h->rank = rank; h->world = world;
    // slices start at multiples of 8192 particles: chunks, groups and estimate partials never straddle ranks
    h->xp.world = world; h->xp.rank = rank;
    for (int r = 0; r <= kMaxPeers; ++r) {
        long long b = r >= world ? h->n : (h->n * r / world) / kSliceAlign * kSliceAlign;
        h->xp.lo[r] = b;
    }
    h->lo = h->xp.lo[rank];
    h->hi = h->xp.lo[rank + 1];
    h->stats.local_particles = h->hi - h->lo;
    CK(cudaMalloc((void**)&h->barrier_word, sizeof(int)));
    CK(cudaMemset(h->barrier_word, 0, sizeof(int)));
    // Map every peer's pose block and exchange block (CUDA IPC over NVLink): pose slices are pushed by the copy engines,
    // the sharded stages exchange their maps and partials by peer stores.  MCL_NO_PEER_PUSH keeps the poses on NCCL
    // all-gathers (A/B); the exchange blocks are required.
    h->peer_pose_block.assign(world, nullptr);
    h->peer_pose_block[rank] = h->pose_block;
    for (int r = 0; r < kMaxPeers; ++r) h->xp.base[r] = nullptr;
    h->xp.base[rank] = h->xblock;
    int ok = world > 1 ? 1 : 0;
    unsigned char* handles_dev = nullptr;
    const size_t hsz = 2 * sizeof(cudaIpcMemHandle_t);          // per rank: pose block, exchange block
    std::vector<cudaIpcMemHandle_t> handles(2 * (size_t)world);
    CK(cudaMalloc((void**)&handles_dev, hsz * world));
    if (ok && cudaIpcGetMemHandle(&handles[2 * rank], h->pose_block) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ok && cudaIpcGetMemHandle(&handles[2 * rank + 1], h->xblock) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    CK(cudaMemcpyAsync(handles_dev + hsz * rank, &handles[2 * rank], hsz, cudaMemcpyHostToDevice, h->stream));
    if (ncclAllGather(handles_dev + hsz * rank, handles_dev, hsz, ncclChar, h->comm, h->stream) != ncclSuccess)
        return fail(h, MCL_ERR_COMM, "NCCL handle exchange failed");
    CK(cudaMemcpyAsync(handles.data(), handles_dev, hsz * world, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < world && ok; ++r) {
        if (r == rank) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, handles[2 * r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        h->peer_pose_block[r] = (float*)p;
        if (cudaIpcOpenMemHandle(&p, handles[2 * r + 1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        h->xp.base[r] = (unsigned char*)p;
    }
    CK(cudaMemcpyAsync(h->barrier_word, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (ncclAllReduce(h->barrier_word, h->barrier_word, 1, ncclInt, ncclMin, h->comm, h->stream) != ncclSuccess)
        return fail(h, MCL_ERR_COMM, "NCCL agreement failed");
    CK(cudaMemcpyAsync(&ok, h->barrier_word, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(handles_dev);
    if (world > 1 && !ok)
        return fail(h, MCL_ERR_COMM, "CUDA IPC mapping of the peers' exchange blocks failed (all ranks must be GPUs of one NVLink node)");
    h->xpeer = world > 1;
    h->peer_push = ok != 0 && !std::getenv("MCL_NO_PEER_PUSH");
    if (h->peer_push) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_action_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_push_done, cudaEventDisableTiming));
    }
    h->stats.peer_push = h->peer_push ? 1 : 0;
    return MCL_OK;
#else
    return fail(h, MCL_ERR_COMM, "library built without NCCL");
#endif
}

// ---- map -----------------------------------------------------------------------------------------------------------
int mcl_set_map(mcl_engine* h, const int8_t* cells, int width, int height, float ox, float oy, float mpc, float cpm)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!cells || width <= 0 || height <= 0) return fail(h, MCL_ERR_INVALID, "bad map");
    CK(cudaSetDevice(h->device));
    const int pitch = (width + 15) & ~15;     // 16-byte rows: aligned word loads for the tile stager (and TMA-ready)
    const int cpitch = (width + 2 * kTabApron + 15) & ~15;
    const size_t crows = (size_t)height + 2 * kTabApron;
    if (!h->map || h->grid.pitch != pitch || h->grid.height != height || h->cpitch != cpitch) {
        if (h->map) {
            CK(cudaStreamSynchronize(h->stream));
            cudaFree(h->map); cudaFree(h->map_fast); cudaFree(h->map_lf); cudaFree(h->dt_steps);
            cudaFree(h->map_cls); cudaFree(h->map_pack);
            h->map = h->map_fast = h->map_lf = nullptr; h->dt_steps = nullptr;
            h->map_cls = nullptr; h->map_pack = nullptr;
        }
        CK(cudaMalloc((void**)&h->map, (size_t)pitch * height + 16));
        CK(cudaMalloc((void**)&h->map_fast, (size_t)pitch * height + 16));
        CK(cudaMalloc((void**)&h->map_lf, (size_t)pitch * height + 16));
        CK(cudaMalloc((void**)&h->map_cls, (size_t)cpitch * crows));
        CK(cudaMalloc((void**)&h->map_pack, sizeof(unsigned long long) * (size_t)cpitch * crows));
        h->cpitch = cpitch;
    }
    CK(cudaMemsetAsync(h->map_cls, 0, (size_t)cpitch * crows, h->stream));
    if (h->dt_steps && (h->grid.width != width || h->grid.height != height)) { cudaFree(h->dt_steps); h->dt_steps = nullptr; }
    CK(cudaMemsetAsync(h->map_lf, 0, (size_t)pitch * height + 16, h->stream));
    CK(cudaMemsetAsync(h->map, 0, (size_t)pitch * height + 16, h->stream));
    CK(cudaMemsetAsync(h->map_fast, 0, (size_t)pitch * height + 16, h->stream));
    CK(cudaMemcpy2DAsync(h->map, pitch, cells, width, width, height, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));     // cells is pageable caller memory: finish before returning
    h->grid.cells = h->map; h->grid.width = width; h->grid.height = height; h->grid.pitch = pitch;
    h->grid.origin_x = ox; h->grid.origin_y = oy; h->grid.cells_per_meter = cpm;
    h->meters_per_cell = mpc;
    h->have_map = true;
    { int rc = refresh_fast_map(h, 0, 0, width, height); if (rc) return rc; }
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

int mcl_update_map_rect(mcl_engine* h, int x0, int y0, int w, int hgt, const int8_t* src, int src_stride)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map has not been called");
    if (!src || w <= 0 || hgt <= 0 || x0 < 0 || y0 < 0 || x0 + w > h->grid.width || y0 + hgt > h->grid.height ||
        src_stride < w)
        return fail(h, MCL_ERR_INVALID, "bad map rectangle");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy2DAsync(h->map + (size_t)y0 * h->grid.pitch + x0, h->grid.pitch, src, src_stride, w, hgt,
                         cudaMemcpyHostToDevice, h->stream));
    { int rc = refresh_fast_map(h, x0, y0, w, hgt); if (rc) return rc; }
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

int mcl_read_map_rect(mcl_engine* h, int x0, int y0, int w, int hgt, int8_t* dst, int dst_stride)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map has not been called");
    if (!dst || w <= 0 || hgt <= 0 || x0 < 0 || y0 < 0 || x0 + w > h->grid.width || y0 + hgt > h->grid.height ||
        dst_stride < w)
        return fail(h, MCL_ERR_INVALID, "bad map rectangle");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy2DAsync(dst, dst_stride, h->map + (size_t)y0 * h->grid.pitch + x0, h->grid.pitch, w, hgt,
                         cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

// ObstacleDistanceGrid::setDistances (planning/obstacle_distance_grid.cpp:73-188) of the mirror.
int mcl_distance_grid(mcl_engine* h, float* out)
{
    if (!h || !out) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map has not been called");
    CK(cudaSetDevice(h->device));
    const int W = h->grid.width, H = h->grid.height;
    int rc = run_distance_transform(h, 0);          // the reference's sources: log-odds >= 0 (occupied or unknown)
    if (rc) return rc;
    h->lf_dirty = true;                             // the scratch no longer holds the likelihood field's steps
    // d_k = fl(d_{k-1} + 0.1f): the reference adds 0.1f once per step (:179)
    std::vector<float> table((size_t)W + H + 2);
    table[0] = 0.0f;
    for (size_t k = 1; k < table.size(); ++k) table[k] = table[k - 1] + 0.1f;
    const size_t cells = (size_t)W * H;
    rc = ensure_staging(h, sizeof(float) * (cells + table.size()));
    if (rc) return rc;
    float* dtab = (float*)h->staging + cells;
    CK(cudaMemcpyAsync(dtab, table.data(), sizeof(float) * table.size(), cudaMemcpyHostToDevice, h->stream));
    dt_to_float_kernel<<<grid_for(h, (long long)cells, 256), 256, 0, h->stream>>>(h->dt_steps, (long long)cells, dtab,
                                                                                (int)table.size(), (float*)h->staging);
    CKL(h);
    CK(cudaMemcpyAsync(out, h->staging, sizeof(float) * cells, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

// Mapping::updateMap (mapping.cpp:17-40) applied to the device mirror.
int mcl_map_update(mcl_engine* h, const mcl_pose_t* previous_pose, const mcl_pose_t* pose, int initialized,
                   const float* ranges, const float* thetas, const int64_t* times, int num_ranges,
                   float max_laser_distance, int hit_odds, int miss_odds, int* rect_xywh_out)
{
    if (!h || !previous_pose || !pose) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map has not been called");
    if (num_ranges < 0 || (num_ranges > 0 && (!ranges || !thetas || !times))) return fail(h, MCL_ERR_INVALID, "bad scan arrays");
    if (num_ranges > 65535)
        return fail(h, MCL_ERR_INVALID, "map update: at most 65535 beams per scan (16-bit per-cell visit counters)");
    if (hit_odds < 0 || hit_odds > 127 || miss_odds < 0 || miss_odds > 127)
        return fail(h, MCL_ERR_INVALID, "hit/miss odds must be in [0, 127] (the parallel form relies on one-signed saturating adds)");
    if (rect_xywh_out) rect_xywh_out[0] = rect_xywh_out[1] = rect_xywh_out[2] = rect_xywh_out[3] = 0;
    // mapping.cpp:19-21: the first call latches the pose and changes no cell (increase/decreaseCellOdds test initialized_)
    if (!initialized || num_ranges == 0) return MCL_OK;
    if (!std::isfinite(pose->x) || !std::isfinite(pose->y) || !std::isfinite(pose->theta) ||
        !std::isfinite(previous_pose->x) || !std::isfinite(previous_pose->y) || !std::isfinite(previous_pose->theta))
        return fail(h, MCL_ERR_INVALID, "non-finite pose");
    CK(cudaSetDevice(h->device));
    if (num_ranges > h->map_beams_cap) {
        if (h->map_beams) cudaFree(h->map_beams);
        if (h->map_beams_host) cudaFreeHost(h->map_beams_host);
        h->map_beams = nullptr; h->map_beams_host = nullptr;
        const int cap = std::max(num_ranges, 1024);
        CK(cudaMalloc((void**)&h->map_beams, sizeof(Beam) * cap));
        CK(cudaMallocHost((void**)&h->map_beams_host, sizeof(Beam) * cap));
        h->map_beams_cap = cap;
    }
    if (!h->map_flag) {
        CK(cudaMalloc((void**)&h->map_flag, sizeof(int)));
        CK(cudaMallocHost((void**)&h->map_flag_host, sizeof(int)));
    }
    // MovingLaserScan(scan, previousPose_, pose): beams with range > 0.15 m, ratio between the two poses' utimes
    const bool interp = previous_pose->utime != pose->utime;              // interpolation.hpp:29-34
    const double denom = (double)(pose->utime - previous_pose->utime);
    int k = 0;
    float reach_m = 0.0f;
    for (int i = 0; i < num_ranges; ++i) {
        if (ranges[i] > h->params.min_range) {                            // moving_laser_scan.cpp:24
            Beam b;
            b.range = ranges[i];
            b.theta = thetas[i];
            b.ratio = interp ? (double)(times[i] - previous_pose->utime) / denom : 1.0;
            h->map_beams_host[k++] = b;
            if (ranges[i] <= max_laser_distance) reach_m = std::max(reach_m, ranges[i]);
        }
    }
    if (k == 0 || reach_m == 0.0f) return MCL_OK;
    // count window: both poses (and every interpolated / extrapolated origin between them) +- reach, clipped to the grid
    double rlo = 0.0, rhi = 1.0;
    for (int i = 0; i < k; ++i) { rlo = std::min(rlo, h->map_beams_host[i].ratio); rhi = std::max(rhi, h->map_beams_host[i].ratio); }
    const double cpm = h->grid.cells_per_meter;
    auto cellx = [&](double x) { return (x - h->grid.origin_x) * cpm; };
    auto celly = [&](double y) { return (y - h->grid.origin_y) * cpm; };
    const double dxm = (double)pose->x - previous_pose->x, dym = (double)pose->y - previous_pose->y;
    const double xs[2] = {previous_pose->x + dxm * rlo, previous_pose->x + dxm * rhi};
    const double ys[2] = {previous_pose->y + dym * rlo, previous_pose->y + dym * rhi};
    const double reach = (double)reach_m * cpm + 3.0;
    long long x0 = (long long)std::floor(std::min(cellx(xs[0]), cellx(xs[1])) - reach);
    long long y0 = (long long)std::floor(std::min(celly(ys[0]), celly(ys[1])) - reach);
    long long x1 = (long long)std::ceil(std::max(cellx(xs[0]), cellx(xs[1])) + reach);
    long long y1 = (long long)std::ceil(std::max(celly(ys[0]), celly(ys[1])) + reach);
    x0 = std::max<long long>(x0, 0); y0 = std::max<long long>(y0, 0);
    x1 = std::min<long long>(x1, h->grid.width - 1); y1 = std::min<long long>(y1, h->grid.height - 1);
    if (x1 < x0 || y1 < y0) return MCL_OK;                                // every touched cell is outside the grid
    const int ww = (int)(x1 - x0 + 1), wh = (int)(y1 - y0 + 1);
    const size_t need = (size_t)ww * wh;
    if (need > h->map_counts_cap) {
        if (h->map_counts) cudaFree(h->map_counts);
        h->map_counts = nullptr; h->map_counts_cap = 0;
        CK(cudaMalloc((void**)&h->map_counts, need * sizeof(uint32_t)));
        h->map_counts_cap = need;
    }
    CK(cudaMemcpyAsync(h->map_beams, h->map_beams_host, sizeof(Beam) * k, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->map_counts, 0, need * sizeof(uint32_t), h->stream));
    CK(cudaMemsetAsync(h->map_flag, 0, sizeof(int), h->stream));
    MapUpdateArgs a{};
    a.beams = h->map_beams; a.num_beams = k;
    a.xa = pose->x; a.ya = pose->y; a.tha = pose->theta;
    a.xb = previous_pose->x; a.yb = previous_pose->y; a.thb = previous_pose->theta;
    a.max_laser_distance = max_laser_distance;
    a.grid = h->grid;
    a.wx0 = (int)x0; a.wy0 = (int)y0; a.ww = ww; a.wh = wh;
    a.counts = h->map_counts;
    a.error_flag = h->map_flag;
    const int blocks = (k + 127) / 128;
    if (interp) map_count_kernel<true><<<blocks, 128, 0, h->stream>>>(a);
    else map_count_kernel<false><<<blocks, 128, 0, h->stream>>>(a);
    CKL(h);
    CK(cudaMemcpyAsync(h->map_flag_host, h->map_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (*h->map_flag_host != 0)       // nothing has been written to the map yet
        return fail(h, MCL_ERR_INVALID, *h->map_flag_host == 1 ? "map update: ray coordinates out of range"
                                                               : "map update: a ray left the count window");
    map_apply_kernel<<<grid_for(h, (long long)need, 256), 256, 0, h->stream>>>(h->map, h->grid.pitch, a.wx0, a.wy0, ww, wh,
                                                                              h->map_counts, hit_odds, miss_odds);
    CKL(h);
    { int rc = refresh_fast_map(h, a.wx0, a.wy0, ww, wh); if (rc) return rc; }
    CK(cudaStreamSynchronize(h->stream));
    if (rect_xywh_out) { rect_xywh_out[0] = a.wx0; rect_xywh_out[1] = a.wy0; rect_xywh_out[2] = ww; rect_xywh_out[3] = wh; }
    return MCL_OK;
}

// ---- particles -----------------------------------------------------------------------------------------------------
int mcl_init_at_pose(mcl_engine* h, float x, float y, float theta, int64_t utime, uint64_t seed)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    CK(cudaSetDevice(h->device));
    h->seed = seed;
    h->update_no = 0;
    const PoseSoA& p = h->pose[h->cur];
    const PoseSoA& q = h->parent[h->cur];
    init_at_pose_kernel<<<grid_for(h, h->n, 256), 256, 0, h->stream>>>(p.x, p.y, p.th, q.x, q.y, q.th, h->n, x, y,
                                                                      theta, h->params.init_std, seed);
    CKL(h);
    fill_kernel<<<grid_for(h, h->n, 256), 256, 0, h->stream>>>(h->weight[h->wcur], h->n, 1.0 / (double)h->n);
    CKL(h);
    h->pose_utime = h->parent_utime = utime;     // particle_filter.cpp:29-30
    h->tab_ok = false;
    h->tab_hint_pending = false;        // (a plan of the previous cloud says nothing about this one)
    h->tab_batch = kTabBatch;
    h->tab_excluded = 0;
    h->cull_interval = 1; h->cull_wait = 0;
    h->have_particles = true;
    h->have_scores = false;
    h->last_estimate.x = x; h->last_estimate.y = y; h->last_estimate.theta = theta; h->last_estimate.utime = utime;
    return MCL_OK;
}

int mcl_init_uniform(mcl_engine* h, int64_t utime, uint64_t seed)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map must precede mcl_init_uniform");
    CK(cudaSetDevice(h->device));
    h->seed = seed;
    h->update_no = 0;
    const PoseSoA& p = h->pose[h->cur];
    const PoseSoA& q = h->parent[h->cur];
    // blocks of about kBatchParticles particles each, as square as the map allows
    const double target = std::max(1.0, (double)h->n / (double)kBatchParticles);
    const double side = std::sqrt((double)h->grid.width * (double)h->grid.height / target);
    const int nbx = (int)std::max(1.0, std::floor((double)h->grid.width / side));
    const int nby = (int)std::max(1.0, std::floor((double)h->grid.height / side));
    init_uniform_kernel<<<grid_for(h, h->n, 256), 256, 0, h->stream>>>(
        p.x, p.y, p.th, q.x, q.y, q.th, h->n, h->grid.origin_x, h->grid.origin_y,
        h->grid.width * h->meters_per_cell, h->grid.height * h->meters_per_cell, seed, nbx, nby);
    CKL(h);
    fill_kernel<<<grid_for(h, h->n, 256), 256, 0, h->stream>>>(h->weight[h->wcur], h->n, 1.0 / (double)h->n);
    CKL(h);
    h->pose_utime = h->parent_utime = utime;
    h->tab_ok = false;
    h->tab_hint_pending = false;        // (a plan of the previous cloud says nothing about this one)
    h->tab_batch = kTabBatch;
    h->tab_excluded = 0;
    h->cull_interval = 1; h->cull_wait = 0;
    h->have_particles = true;
    h->have_scores = false;
    return MCL_OK;
}

int mcl_import_particles(mcl_engine* h, const mcl_particle_t* aos, int64_t n)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!aos || n != h->n) return fail(h, MCL_ERR_INVALID, "import needs exactly num_particles particles");
    for (int64_t i = 1; i < n; ++i)
        if (aos[i].pose.utime != aos[0].pose.utime || aos[i].parent_pose.utime != aos[0].parent_pose.utime)
            return fail(h, MCL_ERR_INVALID, "particles must share one pose.utime and one parent_pose.utime (particle %lld differs)",
                        (long long)i);
    CK(cudaSetDevice(h->device));
    int rc = ensure_staging(h, sizeof(mcl_particle_t) * (size_t)n);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->staging, aos, sizeof(mcl_particle_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    const PoseSoA& p = h->pose[h->cur];
    const PoseSoA& q = h->parent[h->cur];
    aos_to_soa_kernel<<<grid_for(h, n, 256), 256, 0, h->stream>>>((const AosParticle*)h->staging, n, p.x, p.y, p.th,
                                                                 q.x, q.y, q.th, h->weight[h->wcur]);
    CKL(h);
    CK(cudaStreamSynchronize(h->stream));
    h->pose_utime = aos[0].pose.utime;
    h->parent_utime = aos[0].parent_pose.utime;
    h->tab_ok = false;
    h->tab_hint_pending = false;        // (a plan of the previous cloud says nothing about this one)
    h->tab_batch = kTabBatch;
    h->tab_excluded = 0;
    h->cull_interval = 1; h->cull_wait = 0;
    h->have_particles = true;
    h->have_scores = false;
    return MCL_OK;
}

int mcl_export_particles(mcl_engine* h, mcl_particle_t* aos, int64_t max_n, int64_t stride, int64_t* n_out)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!aos || max_n < 0 || stride < 1) return fail(h, MCL_ERR_INVALID, "bad export arguments");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    CK(cudaSetDevice(h->device));
    const long long count = std::min<long long>(max_n, (h->n + stride - 1) / stride);
    if (h->world > 1) {   // parent poses live only on the owning rank: collective call in multi-GPU mode
        int rcj = join_pushes(h);
        if (rcj) return rcj;
        const PoseSoA& q = h->parent[h->cur];
        for (float* arr : {q.x, q.y, q.th}) {
            int rc = exchange_slices(h, arr, sizeof(float));
            if (rc) return rc;
        }
        int rc = exchange_slices(h, h->weight[h->wcur], sizeof(double));     // weights too: every rank normalises its own slice
        if (rc) return rc;
    }
    if (count > 0) {
        int rc = ensure_staging(h, sizeof(mcl_particle_t) * (size_t)count);
        if (rc) return rc;
        const PoseSoA& p = h->pose[h->cur];
        const PoseSoA& q = h->parent[h->cur];
        soa_to_aos_kernel<<<grid_for(h, count, 256), 256, 0, h->stream>>>((AosParticle*)h->staging, count, stride,
                                                                         h->pose_utime, h->parent_utime, p.x, p.y,
                                                                         p.th, q.x, q.y, q.th, h->weight[h->wcur]);
        CKL(h);
        CK(cudaMemcpyAsync(aos, h->staging, sizeof(mcl_particle_t) * (size_t)count, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    if (n_out) *n_out = count;
    return MCL_OK;
}

int mcl_export_weighted(mcl_engine* h, mcl_particle_t* aos, int64_t count, double u01, int64_t* n_out)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!aos || count < 1 || count > h->n || !(u01 >= 0.0 && u01 < 1.0)) return fail(h, MCL_ERR_INVALID, "bad export arguments");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    if (h->world > 1) return fail(h, MCL_ERR_STATE, "weighted export runs on single-GPU engines");
    CK(cudaSetDevice(h->device));
    int rc = ensure_staging(h, sizeof(mcl_particle_t) * (size_t)count + 4 * (size_t)count + 64);
    if (rc) return rc;
    int32_t* pick = (int32_t*)((char*)h->staging + ((sizeof(mcl_particle_t) * (size_t)count + 63) & ~(size_t)63));
    // systematic draw of `count` particles over the weights' exact running sum (the filter's own resampling rule,
    // particle_filter.cpp:84-103, with `count` draws instead of N): draw m sits at (u01 + m) / count of the total mass
    rc = run_resample_indices(h, u01 / (double)count, h->wcur, count, 0, count, pick);
    if (rc) return rc;
    const PoseSoA& p = h->pose[h->cur];
    const PoseSoA& q = h->parent[h->cur];
    soa_to_aos_idx_kernel<<<grid_for(h, count, 256), 256, 0, h->stream>>>((AosParticle*)h->staging, count, pick,
                                                                         1.0 / (double)count, h->pose_utime, h->parent_utime,
                                                                         p.x, p.y, p.th, q.x, q.y, q.th);
    CKL(h);
    CK(cudaMemcpyAsync(aos, h->staging, sizeof(mcl_particle_t) * (size_t)count, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n_out) *n_out = count;
    return MCL_OK;
}

// ---- host scalar action model (action_model.cpp:22-75) ---------------------------------------------------------------
void mcl_action_reset(mcl_action_t* a) { std::memset(a, 0, sizeof(*a)); }

static double angle_diff_host(double l, double r)
{
    double d = l - r;
    if (std::fabs(d) > M_PI) d -= (d > 0) ? M_PI * 2 : M_PI * -2;
    return d;
}

int mcl_action_update(mcl_action_t* a, const mcl_pose_t* odom)
{
    if (!a->initialized) { a->previous_odometry = *odom; a->initialized = 1; }
    const mcl_pose_t& prev = a->previous_odometry;
    const float dx = odom->x - prev.x;                                   // float deltas (:29-30)
    const float dy = odom->y - prev.y;
    const float dth = (float)angle_diff_host((double)odom->theta, (double)prev.theta);
    float dir = 1.0f;
    a->rot1 = angle_diff_host((double)atan2f(dy, dx), (double)prev.theta);   // float atan2 (:34)
    a->trans = (double)sqrtf(dx * dx + dy * dy);                         // float sqrt (:35)
    if (std::fabs(a->trans) < 0.0001) {
        a->rot1 = 0.0;
    } else if (std::fabs(a->rot1) > M_PI / 2.0) {                        // backward motion (:40-43)
        a->rot1 = -angle_diff_host(M_PI, a->rot1);
        dir = -1.0f;
    }
    a->trans *= (double)dir;
    a->rot2 = angle_diff_host((double)dth, a->rot1);
    a->moved = ((std::fabs(a->trans) + std::fabs(a->rot2)) < (double)0.00001f) ? 0 : 1;   // (:52)
    a->rot1_std = 0.05; a->trans_std = 0.005; a->rot2_std = 0.05;        // (:64-66)
    a->previous_odometry = *odom;
    return a->moved;
}

// ---- stages ----------------------------------------------------------------------------------------------------------
int mcl_resample(mcl_engine* h, double r, const double* weights, int32_t* indices_out)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    if (h->world > 1) return fail(h, MCL_ERR_STATE, "stand-alone stages run on single-GPU engines; use mcl_update");
    CK(cudaSetDevice(h->device));
    h->launches = 0;
    double* w = h->weight[h->wcur];
    if (weights) CK(cudaMemcpyAsync(w, weights, sizeof(double) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    int rc = run_resample_indices(h, r, h->wcur);
    if (rc) return rc;
    GatherArgs g{};
    const int s = h->cur, d = h->cur ^ 1;
    g.sx = h->pose[s].x; g.sy = h->pose[s].y; g.sth = h->pose[s].th;
    g.spx = h->parent[s].x; g.spy = h->parent[s].y; g.spth = h->parent[s].th;
    g.sw = w;
    g.dx = h->pose[d].x; g.dy = h->pose[d].y; g.dth = h->pose[d].th;
    g.dpx = h->parent[d].x; g.dpy = h->parent[d].y; g.dpth = h->parent[d].th;
    g.dw = h->weight[h->wcur ^ 1];
    g.idx = h->idx; g.n = h->n;
    gather_kernel<<<grid_for(h, h->n, 256), 256, 0, h->stream>>>(g);
    CKL(h);
    h->cur = d;
    h->wcur ^= 1;
    if (indices_out)
        CK(cudaMemcpyAsync(indices_out, h->idx, sizeof(int32_t) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    return read_counters(h);
}

int mcl_apply_action(mcl_engine* h, const mcl_action_t* a, int64_t utime, const float* noise3n)
{
    if (!h || !a) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    CK(cudaSetDevice(h->device));
    const float* nd;
    int rc = upload_noise(h, noise3n, &nd);
    if (rc) return rc;
    rc = run_action(h, a, utime, nd, false);
    if (rc) return rc;
    rc = join_pushes(h);
    if (rc) return rc;
    rc = rank_barrier(h);       // no score exchange follows here: make every rank's pushes visible before returning
    if (rc) return rc;
    ++h->update_no;
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

int mcl_upload_scan(mcl_engine* h, const float* ranges, const float* thetas, const int64_t* times, int nb,
                    int64_t odometry_utime)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));   // the pinned staging buffer may still be in flight
    // utimes the next action step will assign: parent <- current pose.utime, pose <- odometry utime
    const long long t_end = h->params.legacy_equal_utime ? h->pose_utime : odometry_utime;
    return prepare_scan(h, ranges, thetas, times, nb, h->pose_utime, t_end);
}

int mcl_score(mcl_engine* h, const float* ranges, const float* thetas, const int64_t* times, int nb, double* scores_out)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    if (h->world > 1) return fail(h, MCL_ERR_STATE, "stand-alone stages run on single-GPU engines; use mcl_update");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    int rc = prepare_scan(h, ranges, thetas, times, nb, h->parent_utime, h->pose_utime);
    if (rc) return rc;
    h->launches = 0;
    rc = run_score(h);
    if (rc) return rc;
    if (scores_out) {
        rc = ensure_staging(h, sizeof(double) * (size_t)h->n);
        if (rc) return rc;
        score_to_double_kernel<<<grid_for(h, h->n, 256), 256, 0, h->stream>>>(h->score2, (double*)h->staging, h->n);
        CKL(h);
        CK(cudaMemcpyAsync(scores_out, h->staging, sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
    }
    return read_counters(h);
}

int mcl_normalize(mcl_engine* h, double* weights_out)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    CK(cudaSetDevice(h->device));
    h->launches = 0;
    int rc = run_normalize(h);
    if (rc) return rc;
    if (weights_out)
        CK(cudaMemcpyAsync(weights_out, h->weight[h->wcur], sizeof(double) * (size_t)h->n, cudaMemcpyDeviceToHost,
                           h->stream));
    return read_counters(h);
}

int mcl_estimate(mcl_engine* h, mcl_pose_t* pose_out)
{
    if (!h || !pose_out) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    CK(cudaSetDevice(h->device));
    int rc = run_estimate(h);
    if (rc) return rc;
    rc = fetch_estimate(h, h->pose_utime);
    if (rc) return rc;
    *pose_out = h->last_estimate;
    return MCL_OK;
}

// ---- fused -----------------------------------------------------------------------------------------------------------
int mcl_update(mcl_engine* h, const mcl_action_t* a, int64_t odometry_utime, const float* ranges, const float* thetas,
               const int64_t* times, int nb, double r, const float* noise3n, mcl_pose_t* pose_out)
{
    if (!h || !a || !pose_out) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    if (!h->have_map) return fail(h, MCL_ERR_STATE, "mcl_set_map has not been called");
    CK(cudaSetDevice(h->device));
    if (a->moved) {                                             // particle_filter.cpp:43
        const float* nd;
        int rc = upload_noise(h, noise3n, &nd);
        if (rc) return rc;
        // the scan's interpolation ratios depend on the utimes the action step is about to assign
        const long long t_end = h->params.legacy_equal_utime ? h->pose_utime : odometry_utime;
        if (nb < 0 || (nb > 0 && (!ranges || !thetas || !times))) return fail(h, MCL_ERR_INVALID, "bad scan arrays");
        CK(cudaStreamSynchronize(h->stream));       // (the scan staging buffer may still be in flight)
        const HostScan scan{ranges, thetas, times, nb, h->pose_utime, t_end};
        rc = enqueue_update(h, a, odometry_utime, r, nd, &scan);
        if (rc) return rc;
        rc = read_back(h, true, true, odometry_utime);
        if (rc) return rc;
        prof_collect(h);
        float ms = 0;
        float* slots[5] = {&h->stats.ms_resample, &h->stats.ms_action, &h->stats.ms_score, &h->stats.ms_normalize,
                           &h->stats.ms_estimate};
        for (int i = 0; i < 5; ++i) { cudaEventElapsedTime(slots[i], h->ev[i], h->ev[i + 1]); }
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]);
        h->stats.ms_total = ms;
    }
    h->last_estimate.utime = odometry_utime;                    // particle_filter.cpp:50
    *pose_out = h->last_estimate;
    return MCL_OK;
}

int mcl_update_action_only(mcl_engine* h, const mcl_action_t* a, int64_t odometry_utime, const float* noise3n)
{
    if (!h || !a) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    if (!a->moved) return MCL_OK;                               // particle_filter.cpp:57
    return mcl_apply_action(h, a, odometry_utime, noise3n);
}

int mcl_update_enqueue(mcl_engine* h, const mcl_action_t* a, int64_t odometry_utime, double r)
{
    if (!h || !a) return fail(h, MCL_ERR_INVALID, "null argument");
    if (!h->have_particles || !h->have_map || !h->have_scan) return fail(h, MCL_ERR_STATE, "engine not ready");
    if (!a->moved) return MCL_OK;
    CK(cudaSetDevice(h->device));
    if (r < 0.0) {
        // host-side Philox-free draw is fine here: r only needs to be uniform in [0, 1/N)
        uint64_t s = h->seed ^ (0x9E3779B97F4A7C15ull * (h->update_no + 1));
        s ^= s >> 33; s *= 0xff51afd7ed558ccdULL; s ^= s >> 33;
        r = ((double)(s >> 11) * (1.0 / 9007199254740992.0)) / (double)h->n;
    }
    const int rc = enqueue_update(h, a, odometry_utime, r, nullptr);
    h->prof_n = 0;
    return rc;
}

int mcl_read_estimate(mcl_engine* h, mcl_pose_t* pose_out)
{
    if (!h || !pose_out) return fail(h, MCL_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    int rc = fetch_estimate(h, h->pose_utime);
    if (rc) return rc;
    *pose_out = h->last_estimate;
    return MCL_OK;
}

// ---- introspection -----------------------------------------------------------------------------------------------------
int mcl_get_stats(mcl_engine* h, mcl_stats* out)
{
    if (!h || !out) return fail(h, MCL_ERR_INVALID, "null argument");
    *out = h->stats;
    return MCL_OK;
}

int mcl_set_gather_counting(mcl_engine* h, int on)
{
    if (!h) return fail(h, MCL_ERR_INVALID, "null engine");
    h->count_gathers = on != 0;
    return MCL_OK;
}

int mcl_measure_gather_peak(mcl_engine* h, int64_t footprint_bytes, int64_t reads, double* sectors_per_s_out)
{
    if (!h || !sectors_per_s_out || footprint_bytes < 1024 || reads < 1) return fail(h, MCL_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(h->device));
    unsigned long long pow2 = 1024;
    while ((long long)(pow2 << 1) <= footprint_bytes) pow2 <<= 1;
    int8_t* buf = nullptr;
    CK(cudaMalloc((void**)&buf, pow2));
    CK(cudaMemsetAsync(buf, 1, pow2, h->stream));
    const int blocks = h->sm_count * 8, threads = 256;
    const long long per_thread = std::max<long long>(1, reads / ((long long)blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather_peak_kernel<<<blocks, threads, 0, h->stream>>>(buf, pow2 - 1, per_thread / 4 + 1, h->overruns);   // warm-up
    CK(cudaEventRecord(e0, h->stream));
    gather_peak_kernel<<<blocks, threads, 0, h->stream>>>(buf, pow2 - 1, per_thread, h->overruns);
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    *sectors_per_s_out = (double)per_thread * blocks * threads / ((double)ms * 1e-3);
    return MCL_OK;
}

int mcl_debug_fast_margin(mcl_engine* h, double* max_dev_endpoint, double* max_dev_extended, double* eps_out)
{
    if (!h || !max_dev_endpoint || !max_dev_extended || !eps_out) return fail(h, MCL_ERR_INVALID, "bad arguments");
    if (!h->have_map || !h->have_particles || !h->have_scan) return fail(h, MCL_ERR_STATE, "needs a map, particles and a scan");
    CK(cudaSetDevice(h->device));
    ScoreArgs a{};
    const PoseSoA& p = h->pose[h->cur];
    const PoseSoA& q = h->parent[h->cur];
    a.x = p.x; a.y = p.y; a.th = p.th; a.px = q.x; a.py = q.y; a.pth = q.th;
    a.lo = h->lo; a.hi = h->hi;
    a.beams = h->beams; a.num_beams = h->num_beams;
    a.grid = h->grid;
    a.fast = fast_plan(h, 0, 0, h->grid.width, h->grid.height, h->grid.pitch, std::max(h->grid.width, h->grid.height));
    *eps_out = a.fast.enabled ? h->stats_eps : 0.0;
    *max_dev_endpoint = *max_dev_extended = 0.0;
    if (!a.fast.enabled || h->num_beams == 0) return MCL_OK;
    { int rc = ensure_staging(h, 8); if (rc) return rc; }
    CK(cudaMemsetAsync(h->staging, 0, 8, h->stream));
    const size_t smem = (size_t)h->num_beams * sizeof(Beam);
    if (h->scan_interp) {
        CK(cudaFuncSetAttribute(fast_margin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fast_margin_kernel<true><<<grid_for(h, h->hi - h->lo, 128), 128, smem, h->stream>>>(a, (unsigned*)h->staging);
    } else {
        CK(cudaFuncSetAttribute(fast_margin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fast_margin_kernel<false><<<grid_for(h, h->hi - h->lo, 128), 128, smem, h->stream>>>(a, (unsigned*)h->staging);
    }
    CKL(h);
    float dev[2];
    CK(cudaMemcpyAsync(dev, h->staging, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *max_dev_endpoint = dev[0];
    *max_dev_extended = dev[1];
    return MCL_OK;
}

int mcl_debug_fast_trig_error(mcl_engine* h, float lo, float hi, double* max_sin_err, double* max_cos_err)
{
    if (!h || !max_sin_err || !max_cos_err || !(lo <= hi)) return fail(h, MCL_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(h->device));
    { int rc = ensure_staging(h, 16); if (rc) return rc; }
    CK(cudaMemsetAsync(h->staging, 0, 16, h->stream));
    fast_trig_error_kernel<<<h->sm_count * 16, 256, 0, h->stream>>>(lo, hi, (unsigned long long*)h->staging);
    CKL(h);
    unsigned long long bits[2];
    CK(cudaMemcpyAsync(bits, h->staging, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::memcpy(max_sin_err, &bits[0], 8);
    std::memcpy(max_cos_err, &bits[1], 8);
    return MCL_OK;
}

int mcl_debug_digest(mcl_engine* h, uint64_t* digest4_out)
{
    if (!h || !digest4_out) return fail(h, MCL_ERR_INVALID, "bad arguments");
    if (!h->have_particles) return fail(h, MCL_ERR_STATE, "no particles");
    CK(cudaSetDevice(h->device));
    { int rc = ensure_staging(h, 32); if (rc) return rc; }
    CK(cudaMemsetAsync(h->staging, 0, 32, h->stream));
    const PoseSoA& p = h->pose[h->cur];
    if (h->hi > h->lo) {
        xdigest_kernel<<<grid_for(h, h->hi - h->lo, 256), 256, 0, h->stream>>>(h->idx, h->score2, h->weight[h->wcur], p.x, p.y, p.th,
                                                                             h->lo, h->hi, (unsigned long long*)h->staging);
        CKL(h);
    }
    CK(cudaMemcpyAsync(digest4_out, h->staging, 32, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

int mcl_debug_sincosf(mcl_engine* h, const float* x, int64_t n, float* s, float* c)
{
    if (!h || !x || !s || !c || n < 1) return fail(h, MCL_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(h->device));
    int rc = ensure_staging(h, sizeof(float) * 3 * (size_t)n);
    if (rc) return rc;
    float* dx = (float*)h->staging;
    CK(cudaMemcpyAsync(dx, x, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    debug_sincosf_kernel<<<grid_for(h, n, 256), 256, 0, h->stream>>>(dx, n, dx + n, dx + 2 * n);
    CKL(h);
    CK(cudaMemcpyAsync(s, dx + n, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(c, dx + 2 * n, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCL_OK;
}

}  // extern "C"
