// Kernels of the MCL update for sm_100a.  See DESIGN.md for the data layout and the per-kernel rooflines.
#pragma once
#include "mcl_device.cuh"
#include <type_traits>

#ifndef MCL_BEAM_UNROLL
#define MCL_BEAM_UNROLL 2
#endif
#define MCL_PRAGMA_(x) _Pragma(#x)
#define MCL_UNROLL(n) MCL_PRAGMA_(unroll n)
#ifndef MCL_SCORE_MIN_CTAS
#define MCL_SCORE_MIN_CTAS 2
#endif
#ifndef MCL_SCORE_THREADS
#define MCL_SCORE_THREADS 256
#endif

namespace mcl {

// =================================================================================================================
// K3  sensor model.  G lanes share one particle (G = 32 is "a warp per particle, a lane per beam"; smaller G keeps
// more particles in flight per warp).  Beams are staged in shared memory once per CTA; the particle's ray base lives in
// registers; each lane scores beams g, g+G, ... and the per-particle half-unit score is reduced with warp shuffles.
// TILE: the map window [tile_x0, tile_x0+tile_w) x [tile_y0, tile_y0+tile_h) is staged in shared memory (zero-filled
// outside the grid, which is exactly OccupancyGrid::logOdds' out-of-grid value); reads outside the window fall back to
// the global mirror, so the result never depends on the window choice.
// =================================================================================================================
// Per-batch map windows (global localisation: the cloud as a whole does not fit one shared-memory tile, but consecutive
// particles are spatial neighbours -- mcl_init_uniform lays them out block by block and systematic resampling keeps
// the order -- so every batch of kBatchParticles particles gets its own window, staged by the CTA that scores it).
constexpr int kMaxPeers = 8;
constexpr int kBatchParticles = 1024;
constexpr int kBatchFastThreads = 1024;      // pass 1: one particle per thread, one CTA per SM
constexpr int kBatchExactThreads = 512;      // exact-only kernel
constexpr int kBatchDefThreads = 512;        // pass 2: 16 warps x 2 rounds of 32 particles
struct __align__(16) BatchWindow { int x0, y0, w, h, pitch, bytes, pad0, pad1; };

struct ScoreArgs {
    const float *x, *y, *th;        // pose (SoA)
    const float *px, *py, *pth;     // parent pose
    int32_t* score2;                // out: per-particle score in half units
    long long lo, hi;               // particle slice scored by this launch
    const Beam* beams;
    int num_beams;
    DevGrid grid;
    int tile_x0, tile_y0, tile_w, tile_h, tile_pitch;   // TILE only
    unsigned long long* gather_counter;                  // COUNT only
    FastPlan fast;                                       // two-pass path only
    const int8_t* fast_cells;                            // two-pass path: the fast pass's view of the map (same pitch)
    uint32_t* masks;                                     // two-pass path: [word][virtual lane] uncertain-beam bits
    unsigned long long* deferred_counter;                // two-pass path: evaluations re-done by the exact pass
    const BatchWindow* windows;                          // BATCH only: one map window per kBatchParticles particles
    long long num_batches;
    // Multi-GPU: the kernel that produces a particle's FINAL score stores it straight into every rank's score array
    // over NVLink (CUDA-IPC-mapped peer memory, own array included), so no all-gather of scores follows the sensor
    // stage -- only a 4-byte barrier.  num_peers == 0: single GPU (or NCCL fallback), scores go to score2 only.
    int num_peers;
    int32_t* peer_score[kMaxPeers];
};

// Stages a map window in shared memory (4-byte granules; pitch and x0 are multiples of 4, the mirror's pitch is a
// multiple of 16; mirror rows are zero-padded to the pitch; outside the grid reads as 0).
__device__ __forceinline__ void stage_window(const DevGrid& grid, int x0, int y0, int hgt, int pitch, int8_t* stile)
{
    const int words_per_row = pitch >> 2;
    const int total = words_per_row * hgt;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int ty = i / words_per_row, tw = i - ty * words_per_row;
        const int gx = x0 + (tw << 2), gy = y0 + ty;
        uint32_t v = 0;
        if ((unsigned)gy < (unsigned)grid.height && gx >= 0 && gx < grid.pitch)
            v = __ldg(reinterpret_cast<const uint32_t*>(grid.cells + (size_t)gy * grid.pitch + gx));
        reinterpret_cast<uint32_t*>(stile)[i] = v;
    }
}
__device__ __forceinline__ void stage_tile(const ScoreArgs& a, int8_t* stile)
{
    stage_window(a.grid, a.tile_x0, a.tile_y0, a.tile_h, a.tile_pitch, stile);
}

__device__ __forceinline__ Window make_window(const ScoreArgs& a, const int8_t* stile, bool tile)
{
    Window win;
    if (tile) {
        win.base = stile; win.x0 = a.tile_x0; win.y0 = a.tile_y0; win.w = a.tile_w; win.h = a.tile_h;
        win.pitch = a.tile_pitch;
    } else {
        win.base = a.grid.cells; win.x0 = 0; win.y0 = 0; win.w = a.grid.width; win.h = a.grid.height;
        win.pitch = a.grid.pitch;
    }
    return win;
}

// The window-dependent fields of the fast plan (the eps-dependent ones are set by the host: mcl_engine.cu fast_plan).
// Everything here is in WINDOW-RELATIVE cells: cell (x0, y0) of the grid is (0, 0).
__host__ __device__ inline void plan_set_window(FastPlan& fp, long long x0, long long y0, long long w, long long hh,
                                                long long pitch)
{
    fp.shift_x = (float)x0; fp.shift_y = (float)y0;
    // certain-interior cells [lc, hc): one cell inside the window, and inside the grid (global cell >= 0)
    const long long lcx = (-x0 > 1 ? -x0 : 1), hcx = w - 1;
    const long long lcy = (-y0 > 1 ? -y0 : 1), hcy = hh - 1;
    const bool empty = hcx <= lcx || hcy <= lcy;
    // The cell is read off the fixed-point bits, which carry the band offset (fp.magic): shrink the box by that much
    // (+ half a fixed-point step, fp.band) so that the cell taken from the bits is inside [lc, hc) whenever the test
    // passes, also for endpoints in the uncertain band just below an integer.
    fp.mid_x = 0.5f * (float)(lcx + hcx); fp.half_x = empty ? -1.0f : 0.5f * (float)(hcx - lcx) - fp.band;
    fp.mid_y = 0.5f * (float)(lcy + hcy); fp.half_y = empty ? -1.0f : 0.5f * (float)(hcy - lcy) - fp.band;
    fp.gmid_x = 0.5f * (float)(fp.grid_w - 1) - (float)x0;
    fp.gmid_y = 0.5f * (float)(fp.grid_h - 1) - (float)y0;
    fp.x2_lo_x = fp.x2_min - (float)x0;
    fp.x2_lo_y = fp.x2_min - (float)y0;
    fp.pitch_f = (float)pitch;
    fp.idx_bias = (int)((unsigned)fp.mbk * (unsigned)pitch + (unsigned)fp.mbk);
    fp.safe_idx = (int)pitch + 1;          // an empty window still has (pitch >= 4) bytes of shared memory behind it
}

// Bounding box of a batch's poses and parents -> its map window, with the same rule as the single-tile path
// (bbox +- (max range + 3 cells), clipped to the grid plus a 2-cell zero margin, 4-byte aligned, odd word pitch).
// summary[0] = largest window in bytes among batches that fit the budget, summary[2] = their largest width/height,
// summary[1] = batches that do not fit
// (they get an empty window: pass 1 defers everything, the exact pass reads the global mirror).
__device__ __forceinline__ int float_order_i(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__global__ void __launch_bounds__(256) batch_window_kernel(const float* x, const float* y, const float* px,
                                                           const float* py, long long lo, long long hi, DevGrid grid,
                                                           double reach, int budget_bytes, BatchWindow* out, int* summary)
{
    __shared__ int red[4][8];
    const long long first = lo + (long long)blockIdx.x * kBatchParticles;
    const long long last = first + kBatchParticles < hi ? first + kBatchParticles : hi;
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
    for (long long i = first + threadIdx.x; i < last; i += blockDim.x) {
        const int a = float_order_i(x[i]), b = float_order_i(y[i]), c = float_order_i(px[i]), d = float_order_i(py[i]);
        mnx = min(mnx, min(a, c)); mxx = max(mxx, max(a, c));
        mny = min(mny, min(b, d)); mxy = max(mxy, max(b, d));
    }
    for (int off = 16; off > 0; off >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, off));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, off));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, off));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, off));
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = mnx; red[1][threadIdx.x >> 5] = mny;
        red[2][threadIdx.x >> 5] = mxx; red[3][threadIdx.x >> 5] = mxy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            mnx = min(mnx, red[0][k]); mny = min(mny, red[1][k]); mxx = max(mxx, red[2][k]); mxy = max(mxy, red[3][k]);
        }
        auto unorder = [](int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); };
        const float fx0 = unorder(mnx), fy0 = unorder(mny), fx1 = unorder(mxx), fy1 = unorder(mxy);
        BatchWindow bw = {0, 0, 0, 0, 4, 0, 0, 0};
        bool fits = false;
        if (isfinite(fx0) && isfinite(fy0) && isfinite(fx1) && isfinite(fy1)) {
            const double cpm = (double)grid.cells_per_meter;
            const double cx0 = floor(((double)fx0 - (double)grid.origin_x) * cpm - reach);
            const double cy0 = floor(((double)fy0 - (double)grid.origin_y) * cpm - reach);
            const double cx1 = ceil(((double)fx1 - (double)grid.origin_x) * cpm + reach);
            const double cy1 = ceil(((double)fy1 - (double)grid.origin_y) * cpm + reach);
            long long x0 = (long long)fmax(cx0, -2.0), y0 = (long long)fmax(cy0, -2.0);
            const long long x1 = (long long)fmin(cx1, (double)grid.width + 1);
            const long long y1 = (long long)fmin(cy1, (double)grid.height + 1);
            if (x1 >= x0 && y1 >= y0) {
                x0 = (x0 >= 0) ? (x0 & ~3ll) : -(((-x0) + 3) & ~3ll);
                const long long tw = x1 - x0 + 1, th = y1 - y0 + 1;
                long long pitch = (tw + 3) & ~3ll;
                if (((pitch >> 2) & 1) == 0) pitch += 4;
                const long long bytes = pitch * th;
                if (bytes <= (long long)budget_bytes) {
                    bw.x0 = (int)x0; bw.y0 = (int)y0; bw.w = (int)tw; bw.h = (int)th; bw.pitch = (int)pitch;
                    bw.bytes = (int)bytes;
                    fits = true;
                }
            }
        }
        out[blockIdx.x] = bw;
        if (fits) { atomicMax(summary + 0, bw.bytes); atomicMax(summary + 2, max(bw.w, bw.h)); }
        else atomicAdd(summary + 1, 1);
    }
}

// Work units of the three sensor kernels.  Single window (BATCH = false): one unit, the CTA strides over the whole
// slice.  Per-batch windows (BATCH = true): CTA c takes batches c, c + gridDim.x, ...; before each it stages that
// batch's window.  unit_range() gives the particle range and stride of the current unit.
template <bool BATCH>
__device__ __forceinline__ void unit_range(const ScoreArgs& a, long long unit, int ppb, long long& first, long long& last,
                                           long long& stride)
{
    if (BATCH) {
        first = a.lo + unit * kBatchParticles;
        last = first + kBatchParticles < a.hi ? first + kBatchParticles : a.hi;
        stride = ppb;
    } else {
        first = a.lo + (long long)blockIdx.x * ppb;
        last = a.hi;
        stride = (long long)gridDim.x * ppb;
    }
}

// ---- the literal restatement for every evaluation (sensor_path = 1, and whenever the fast pass is not applicable) ------
template <int G, bool INTERP, bool TILE, bool COUNT, bool BATCH>
__global__ void __launch_bounds__(BATCH ? kBatchExactThreads : MCL_SCORE_THREADS, BATCH ? 1 : MCL_SCORE_MIN_CTAS)
score_kernel(const ScoreArgs a)
{
    constexpr int T = BATCH ? kBatchExactThreads : MCL_SCORE_THREADS;
    extern __shared__ __align__(16) unsigned char smem[];
    Beam* sbeams = reinterpret_cast<Beam*>(smem);
    int8_t* stile = reinterpret_cast<int8_t*>(smem + (size_t)a.num_beams * sizeof(Beam));

    for (int i = threadIdx.x; i < a.num_beams; i += blockDim.x) sbeams[i] = a.beams[i];
    if (TILE && !BATCH) stage_tile(a, stile);
    __syncthreads();

    Window win = make_window(a, stile, TILE);
    GridConst gc;
    gc.gx = (double)a.grid.origin_x; gc.gy = (double)a.grid.origin_y;
    gc.cpm = a.grid.cells_per_meter; gc.cpm_d = (double)a.grid.cells_per_meter;
    gc.trig = gs_load_consts();

    constexpr int PPB = T / G;                   // particles per CTA per iteration
    const int sub = threadIdx.x % G;
    const int slot = threadIdx.x / G;
    int gathers = 0;
    const long long nunits = BATCH ? a.num_batches : 1;
    for (long long unit = BATCH ? blockIdx.x : 0; unit < nunits; unit += BATCH ? gridDim.x : 1) {
        if (BATCH) {
            const BatchWindow bw = a.windows[unit];
            __syncthreads();
            stage_window(a.grid, bw.x0, bw.y0, bw.h, bw.pitch, stile);
            __syncthreads();
            win.x0 = bw.x0; win.y0 = bw.y0; win.w = bw.w; win.h = bw.h; win.pitch = bw.pitch;
        }
        long long first, last, stride;
        unit_range<BATCH>(a, unit, PPB, first, last, stride);
        for (long long base = first; base < last; base += stride) {
            const long long p = base + slot;
            int acc = 0;
            if (p < last) {
                const RayBase rb = make_ray_base(a.x[p], a.y[p], a.th[p], a.px[p], a.py[p], a.pth[p]);
MCL_UNROLL(MCL_BEAM_UNROLL)
                for (int j = sub; j < a.num_beams; j += G)
                    acc += score_beam<INTERP, TILE, COUNT>(rb, sbeams[j], gc, win, a.grid, gathers);
            }
#pragma unroll
            for (int off = G >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (sub == 0 && p < last) {
                if (a.num_peers > 0) {
                    for (int r = 0; r < a.num_peers; ++r) a.peer_score[r][p] = acc;     // final: to every rank
                } else {
                    a.score2[p] = acc;
                }
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gathers += __shfl_xor_sync(0xffffffffu, gathers, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.gather_counter, (unsigned long long)gathers);
    }
}

// ---- two-pass path, pass 1: certified float evaluation (mcl_device.cuh: score_beam_fast) ------------------------------
// Writes the sum over the certain beams to score2[p] and one bit per uncertain beam to masks[word][virtual lane], where
// virtual lane = (p - lo)*G + sub and a lane's k-th beam is j = sub + k*G.
#ifndef MCL_FAST_THREADS
#define MCL_FAST_THREADS 256
#endif
#ifndef MCL_FAST_MIN_CTAS
#define MCL_FAST_MIN_CTAS 4
#endif
#ifndef MCL_FAST_UNROLL
#define MCL_FAST_UNROLL 2
#endif
template <int G, bool INTERP, bool TILE, bool COUNT, bool BATCH>
__global__ void __launch_bounds__(BATCH ? kBatchFastThreads : MCL_FAST_THREADS, BATCH ? 1 : MCL_FAST_MIN_CTAS)
score_fast_kernel(const ScoreArgs a)
{
    constexpr int T = BATCH ? kBatchFastThreads : MCL_FAST_THREADS;
    extern __shared__ __align__(16) unsigned char smem[];
    FastBeam* sfast = reinterpret_cast<FastBeam*>(smem);
    int8_t* stile = reinterpret_cast<int8_t*>(smem + (size_t)a.num_beams * sizeof(FastBeam));

    for (int i = threadIdx.x; i < a.num_beams; i += blockDim.x) {
        const Beam b = a.beams[i];
        FastBeam f;
        f.ratio = (float)b.ratio; f.theta = b.theta; f.rc = __fmul_rn(b.range, a.grid.cells_per_meter); f.pad = 0.0f;
        sfast[i] = f;
    }
    DevGrid fgrid = a.grid;                 // the derived map: positive cells unchanged, non-positive ones 0 or -1
    fgrid.cells = a.fast_cells;
    if (TILE && !BATCH) stage_window(fgrid, a.tile_x0, a.tile_y0, a.tile_h, a.tile_pitch, stile);
    __syncthreads();

    const int8_t* cells = TILE ? stile : fgrid.cells;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(stile);
    int pitch = TILE ? a.tile_pitch : a.grid.pitch;
    FastPlan fp = a.fast;
    const double gx = (double)a.grid.origin_x, gy = (double)a.grid.origin_y, cpm_d = (double)a.grid.cells_per_meter;
    const int iters = (a.num_beams + G - 1) / G;                 // beams per lane
    const int nwords = (iters + 31) / 32;
    const long long vlanes = (a.hi - a.lo) * G;

    constexpr int PPB = T / G;
    const int sub = threadIdx.x % G;
    const int slot = threadIdx.x / G;
    int gathers = 0;
    const long long nunits = BATCH ? a.num_batches : 1;
    for (long long unit = BATCH ? blockIdx.x : 0; unit < nunits; unit += BATCH ? gridDim.x : 1) {
        if (BATCH) {
            const BatchWindow bw = a.windows[unit];
            __syncthreads();
            stage_window(fgrid, bw.x0, bw.y0, bw.h, bw.pitch, stile);
            __syncthreads();
            plan_set_window(fp, bw.x0, bw.y0, bw.w, bw.h, bw.pitch);
            pitch = bw.pitch;
        }
        long long first, last, stride;
        unit_range<BATCH>(a, unit, PPB, first, last, stride);
        for (long long base = first; base < last; base += stride) {
            const long long p = base + slot;
            int acc = 0;
            if (p < last) {
                const FastBase fb =
                    make_fast_base<INTERP>(a.x[p], a.y[p], a.th[p], a.px[p], a.py[p], a.pth[p], gx, gy, cpm_d, fp);
                uint32_t* mrow = a.masks + ((p - a.lo) * G + sub);
                // the edge tests a warp needs are those of its most exposed particle
                const int edge = __reduce_max_sync(__activemask(), fb.ok ? fb.edge : 0);
                for (int w = 0; w < nwords; ++w) {
                    uint32_t m = 0;
                    const int kend = min(32, iters - w * 32);
                    if (fb.ok) {
                        auto run = [&](auto edge_tag) {
                            constexpr int EDGE = decltype(edge_tag)::value;
                            uint32_t bit = 1u;
MCL_UNROLL(MCL_FAST_UNROLL)
                            for (int k = 0; k < kend; ++k) {
                                const int j = sub + (w * 32 + k) * G;
                                const bool inb = G == 1 || j < a.num_beams;      // G == 1: kend already bounds j
                                int v = 0, g = 0;
                                const bool certain = score_beam_fast<INTERP, TILE, COUNT, EDGE>(
                                    fb, sfast[inb ? j : 0], fp, cells, sbase, pitch, v, g);
                                acc += inb ? v : 0;
                                if (COUNT) gathers += inb ? g : 0;
                                if (inb & !certain) m |= bit;
                                bit += bit;
                            }
                        };
                        if (edge == 0) run(std::integral_constant<int, 0>{});
                        else if (edge == 1) run(std::integral_constant<int, 1>{});
                        else run(std::integral_constant<int, 2>{});
                    } else {
                        for (int k = 0; k < kend; ++k) m |= (uint32_t)(sub + (w * 32 + k) * G < a.num_beams) << k;
                    }
                    mrow[(long long)w * vlanes] = m;
                }
            }
#pragma unroll
            for (int off = G >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (sub == 0 && p < last) a.score2[p] = acc;
        }
    }
    if (COUNT) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gathers += __shfl_xor_sync(0xffffffffu, gathers, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.gather_counter, (unsigned long long)gathers);
    }
}

// ---- two-pass path, pass 2: the literal restatement for the beams pass 1 could not certify ------------------------------
// Warp-level compaction: a warp takes 32 consecutive virtual lanes at a time, turns their mask words into a queue of
// (word, lane, bit) entries in shared memory and evaluates the queue 32 entries at a time, so every lane of the warp
// does useful work whatever the distribution of uncertain beams over particles.  The raw poses of the 32 lanes'
// particles are staged in shared memory (coalesced) and each evaluation rebuilds its RayBase from them; results are
// accumulated with shared-memory integer atomics (exact, order-independent) and added to score2[p].
#ifndef MCL_DEF_THREADS
#define MCL_DEF_THREADS 384
#endif
#ifndef MCL_DEF_MIN_CTAS
#define MCL_DEF_MIN_CTAS 2
#endif
constexpr int kDefQueue = 1024 + 32;          // one word of 32 lanes can add up to 1024 entries to < 32 left over
struct __align__(16) DefParticle { float xa, ya, tha, xb, yb, thb; int acc; int pad; };

__host__ __device__ inline size_t deferred_smem_bytes(int num_beams, int warps)
{
    return (size_t)num_beams * sizeof(Beam) + (size_t)warps * (kDefQueue * sizeof(uint16_t) + 32 * sizeof(DefParticle));
}

template <int G, bool INTERP, bool TILE, bool COUNT, bool BATCH>
__global__ void __launch_bounds__(BATCH ? kBatchDefThreads : MCL_DEF_THREADS, BATCH ? 1 : MCL_DEF_MIN_CTAS)
score_deferred_kernel(const ScoreArgs a)
{
    constexpr int T = BATCH ? kBatchDefThreads : MCL_DEF_THREADS;
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int WARPS = T / 32;
    Beam* sbeams = reinterpret_cast<Beam*>(smem);
    DefParticle* spart_all = reinterpret_cast<DefParticle*>(smem + (size_t)a.num_beams * sizeof(Beam));
    uint16_t* squeue_all = reinterpret_cast<uint16_t*>(spart_all + WARPS * 32);
    int8_t* stile = reinterpret_cast<int8_t*>(smem + deferred_smem_bytes(a.num_beams, WARPS));

    for (int i = threadIdx.x; i < a.num_beams; i += blockDim.x) sbeams[i] = a.beams[i];
    if (TILE && !BATCH) stage_tile(a, stile);
    __syncthreads();

    Window win = make_window(a, stile, TILE);
    GridConst gc;
    gc.gx = (double)a.grid.origin_x; gc.gy = (double)a.grid.origin_y;
    gc.cpm = a.grid.cells_per_meter; gc.cpm_d = (double)a.grid.cells_per_meter;
    gc.trig = gs_load_consts();
    const int iters = (a.num_beams + G - 1) / G;
    const int nwords = (iters + 31) / 32;
    const long long vlanes = (a.hi - a.lo) * G;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DefParticle* spart = spart_all + warp * 32;
    uint16_t* squeue = squeue_all + warp * kDefQueue;
    int gathers = 0;
    unsigned deferred = 0;

    auto evaluate = [&](unsigned e) {
        const int k = e & 31, src = (e >> 5) & 31, w = e >> 10;
        const DefParticle dp = spart[src / G];
        const RayBase rb = make_ray_base(dp.xa, dp.ya, dp.tha, dp.xb, dp.yb, dp.thb);
        const int j = (src % G) + (w * 32 + k) * G;
        const int v = score_beam<INTERP, TILE, COUNT>(rb, sbeams[j], gc, win, a.grid, gathers);
        if (v) atomicAdd(&spart[src / G].acc, v);
        ++deferred;
    };

    // work units: blocks of 32 virtual lanes.  BATCH: the blocks of one batch at a time, behind that batch's window.
    const long long nblocks_all = (vlanes + 31) / 32;
    constexpr long long kBlocksPerBatch = (long long)kBatchParticles * G / 32;
    const long long nunits = BATCH ? a.num_batches : 1;
    for (long long unit = BATCH ? blockIdx.x : 0; unit < nunits; unit += BATCH ? gridDim.x : 1) {
        long long bfirst, blast, bstride;
        if (BATCH) {
            const BatchWindow bw = a.windows[unit];
            __syncthreads();
            stage_window(a.grid, bw.x0, bw.y0, bw.h, bw.pitch, stile);
            __syncthreads();
            win.x0 = bw.x0; win.y0 = bw.y0; win.w = bw.w; win.h = bw.h; win.pitch = bw.pitch;
            bfirst = unit * kBlocksPerBatch + warp;
            blast = (unit + 1) * kBlocksPerBatch < nblocks_all ? (unit + 1) * kBlocksPerBatch : nblocks_all;
            bstride = WARPS;
        } else {
            bfirst = (long long)blockIdx.x * WARPS + warp;
            blast = nblocks_all;
            bstride = (long long)gridDim.x * WARPS;
        }
        for (long long blk = bfirst; blk < blast; blk += bstride) {
            const long long v0 = blk * 32;
            const long long v = v0 + lane;
            // stage the particles of this block (32/G of them)
            if (lane < 32 / G) {
                const long long p = a.lo + v0 / G + lane;
                DefParticle dp;
                dp.acc = 0; dp.pad = 0;
                if (p < a.hi) {
                    dp.xa = a.x[p]; dp.ya = a.y[p]; dp.tha = a.th[p]; dp.xb = a.px[p]; dp.yb = a.py[p]; dp.thb = a.pth[p];
                } else {
                    dp.xa = dp.ya = dp.tha = dp.xb = dp.yb = dp.thb = 0.0f;
                }
                spart[lane] = dp;
            }
            __syncwarp();
            int qn = 0;
            uint32_t mbuf[4];                    // mask words are fetched four at a time: one exposed latency per four
            for (int w = 0; w < nwords; ++w) {
                if ((w & 3) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        mbuf[q] = (v < vlanes && w + q < nwords) ? __ldcs(a.masks + ((long long)(w + q) * vlanes + v)) : 0u;
                }
                uint32_t m = mbuf[0];
                mbuf[0] = mbuf[1]; mbuf[1] = mbuf[2]; mbuf[2] = mbuf[3];
                const int c = __popc(m);
                int incl = c;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += t;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0) continue;
                int pos = qn + incl - c;
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    squeue[pos++] = (uint16_t)((w << 10) | (lane << 5) | k);
                }
                qn += total;
                __syncwarp();
                while (qn >= 32) {
                    qn -= 32;
                    evaluate(squeue[qn + lane]);
                }
                __syncwarp();
            }
            if (lane < qn) evaluate(squeue[lane]);
            __syncwarp();
            if (lane < 32 / G) {
                const long long p = a.lo + v0 / G + lane;
                const int add = spart[lane].acc;
                if (p < a.hi) {
                    if (a.num_peers > 0) {
                        const int fin = a.score2[p] + add;                       // pass 1's partial + this pass: final
                        for (int r = 0; r < a.num_peers; ++r) a.peer_score[r][p] = fin;
                    } else if (add != 0) {
                        a.score2[p] += add;
                    }
                }
            }
            __syncwarp();
        }
    }
    if (COUNT) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gathers += __shfl_xor_sync(0xffffffffu, gathers, off);
        if (lane == 0) atomicAdd(a.gather_counter, (unsigned long long)gathers);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) deferred += __shfl_xor_sync(0xffffffffu, deferred, off);
    if (lane == 0 && deferred) atomicAdd(a.deferred_counter, (unsigned long long)deferred);
}

// =================================================================================================================
// K2  action model (action_model.cpp:78-103), optionally fused with the resampling gather (particle_filter.cpp:100):
// child m reads its parent's pose through src_index (or m itself), writes parent_pose = that pose and the moved pose.
// =================================================================================================================
struct ActionArgs {
    const float *sx, *sy, *sth;        // source pose arrays (global)
    const int32_t* src_index;          // null = identity
    float *dx, *dy, *dth;              // destination pose
    float *dpx, *dpy, *dpth;           // destination parent pose
    long long lo, hi;
    double rot1, trans, rot2, s1, st, s2;
    int moved;
    const float* noise;                // 3 floats per particle (global index) or null
    uint64_t seed;
    uint32_t update_no;
};

__global__ void __launch_bounds__(256) action_kernel(const ActionArgs a)
{
    for (long long m = a.lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; m < a.hi;
         m += (long long)gridDim.x * blockDim.x) {
        const long long src = a.src_index ? (long long)a.src_index[m] : m;
        const float x = a.sx[src], y = a.sy[src], th = a.sth[src];
        float nx = x, ny = y, nth = th;
        if (a.moved) {
            float r1, tr, r2;
            if (a.noise) {
                r1 = a.noise[3 * m + 0]; tr = a.noise[3 * m + 1]; r2 = a.noise[3 * m + 2];
            } else {
                const uint4 w = philox4x32_10(make_uint4((uint32_t)m, (uint32_t)((unsigned long long)m >> 32),
                                                         a.update_no, 0x4d434cu),
                                              make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                float z0, z1, z2, z3;
                philox_normals(w, z0, z1, z2, z3);
                r1 = (float)__dadd_rn(__dmul_rn((double)z0, a.s1), a.rot1);   // (float)(z*sigma + mu), double affine
                tr = (float)__dadd_rn(__dmul_rn((double)z1, a.st), a.trans);
                r2 = (float)__dadd_rn(__dmul_rn((double)z2, a.s2), a.rot2);
            }
            const float h = __fadd_rn(th, r1);                                  // float sum (:88)
            double sn, cs;
            sincos((double)h, &sn, &cs);                                        // unqualified cos/sin -> double
            nx = (float)__dadd_rn((double)x, __dmul_rn((double)tr, cs));
            ny = (float)__dadd_rn((double)y, __dmul_rn((double)tr, sn));
            nth = wrap_to_pi(__fadd_rn(h, r2));                                 // (th + r1) + r2, float (:90)
        }
        a.dpx[m] = x; a.dpy[m] = y; a.dpth[m] = th;                             // parent_pose = sample.pose (:93)
        a.dx[m] = nx; a.dy[m] = ny; a.dth[m] = nth;
    }
}

// Whole-particle gather for the stand-alone resample stage (prior[m] = posterior_[i], particle_filter.cpp:100).
struct GatherArgs {
    const float *sx, *sy, *sth, *spx, *spy, *spth;
    const double* sw;
    float *dx, *dy, *dth, *dpx, *dpy, *dpth;
    double* dw;
    const int32_t* idx;
    long long n;
};
__global__ void gather_kernel(const GatherArgs a)
{
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < a.n; m += (long long)gridDim.x * blockDim.x) {
        const int i = a.idx[m];
        a.dx[m] = a.sx[i]; a.dy[m] = a.sy[i]; a.dth[m] = a.sth[i];
        a.dpx[m] = a.spx[i]; a.dpy[m] = a.spy[i]; a.dpth[m] = a.spth[i];
        a.dw[m] = a.sw[i];
    }
}

// =================================================================================================================
// K4  weights and estimate.
// =================================================================================================================
// particle_filter.cpp:126-133: v = score < floor ? floor : score, from half-unit integer scores.
__global__ void floor_kernel(const int32_t* score2, double* v, long long n, double floor_w)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double s = 0.5 * (double)score2[i];
        v[i] = s < floor_w ? floor_w : s;
    }
}

// weight_mode 1 (extension): scores as log-likelihoods.  score_max_kernel finds the largest half-unit score;
// lse_kernel writes v = exp(beta * (score - max)) <= 1, so the sum cannot overflow and the best particle contributes
// exactly 1 (the max / log-sum-exp normalisation); the sum and the division then go through the same path as mode 0.
__global__ void score_max_kernel(const int32_t* score2, long long n, int* out_max)
{
    int m = (int)0x80000000;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, score2[i]);
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(out_max, m);
}
__global__ void lse_kernel(const int32_t* score2, double* v, long long n, const int* score_max, double half_beta)
{
    const int m = *score_max;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = exp(half_beta * (double)(score2[i] - m));        // half units: beta/2 per unit of score2
}

// particle_filter.cpp:136-138: w /= wSum (IEEE double division, correctly rounded on both sides).
__global__ void divide_kernel(double* w, long long n, const double* wsum, double* ess_acc)
{
    const double s = *wsum;
    double sq = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double q = __ddiv_rn(w[i], s);
        w[i] = q;
        sq += q * q;
    }
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((threadIdx.x & 31) == 0 && ess_acc) atomicAdd(ess_acc, sq);
}

__global__ void fill_kernel(double* w, long long n, double v)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        w[i] = v;
}

// particle_filter.cpp:144-160.  Deterministic two-stage reduction of (sum w*x, sum w*y, sum w*sinf, sum w*cosf) in
// double: fixed chunk -> block mapping, fixed tree, so the estimate does not depend on the GPU count or the run.
constexpr int kEstBlock = 256;
constexpr int kEstChunk = 4096;   // particles per partial
__global__ void __launch_bounds__(kEstBlock) estimate_partial_kernel(const float* x, const float* y, const float* th,
                                                                     const double* w, long long n, double4* partials)
{
    __shared__ double4 red[kEstBlock];
    const long long base = (long long)blockIdx.x * kEstChunk;
    double4 acc = make_double4(0, 0, 0, 0);
    for (int k = threadIdx.x; k < kEstChunk; k += kEstBlock) {
        const long long i = base + k;
        if (i < n) {
            const double wi = w[i];
            float s, c;
            glibc_sincosf(th[i], &s, &c);       // std::sin/std::cos(float) -> sincosf in the reference build
            acc.x = __dadd_rn(acc.x, __dmul_rn(wi, (double)x[i]));
            acc.y = __dadd_rn(acc.y, __dmul_rn(wi, (double)y[i]));
            acc.z = __dadd_rn(acc.z, __dmul_rn(wi, (double)s));
            acc.w = __dadd_rn(acc.w, __dmul_rn(wi, (double)c));
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = kEstBlock / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            double4 o = red[threadIdx.x + off], m = red[threadIdx.x];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[threadIdx.x] = m;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

// out4 = (x, y, theta, unused) as floats; single block.
__global__ void __launch_bounds__(kEstBlock) estimate_final_kernel(const double4* partials, int count, float* out4)
{
    __shared__ double4 red[kEstBlock];
    double4 acc = make_double4(0, 0, 0, 0);
    for (int k = threadIdx.x; k < count; k += kEstBlock) {      // fixed order per thread
        const double4 p = partials[k];
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = kEstBlock / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            double4 o = red[threadIdx.x + off], m = red[threadIdx.x];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[threadIdx.x] = m;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out4[0] = (float)red[0].x;
        out4[1] = (float)red[0].y;
        out4[2] = (float)atan2(red[0].z, red[0].w);
        out4[3] = 0.0f;
    }
}

// =================================================================================================================
// K1  resampling.  The reference compares U_m = r + m/N against c, a LEFT-TO-RIGHT double running sum
// (particle_filter.cpp:93-99).  A parallel double scan rounds differently, so indices would differ.  The kernels below
// reproduce the sequential rounding EXACTLY, in parallel:
//   while c stays inside one binade [2^e, 2^(e+1)) every step is c += (w rounded to a multiple of ulp_e, ties by the
//   parity of c/ulp_e), so over a chunk of elements the map c_in -> c_out is "add D[parity(c_in/ulp)]": two integers.
//   Such maps compose associatively.  Chunks whose running sum may cross a binade edge are walked serially with real
//   double adds.  D[0], D[1] themselves are obtained with real double adds from the two representatives 2^e and
//   2^e + ulp, so hardware rounding does the tie handling.
// Level 1: kL1 elements per chunk;  level 2: kL2 level-1 chunks per group.  Every assumption (binade of the true
// running sum at chunk entry, no overflow out of the binade) is verified against exact values during the serial walk;
// failures take the serial path, so the result never depends on the heuristics.
// =================================================================================================================
constexpr int kL1 = 128;
constexpr int kL2 = 64;
constexpr int kSeqTileChunks = 32;            // level-1 chunks staged per CTA
constexpr int kSeqRowPitch = kL1 + 1;         // doubles; odd pitch => conflict-free column walks

__device__ __forceinline__ int dbl_exp(double v) { return (int)((__double_as_longlong(v) >> 52) & 0x7ff); }

// Stage kSeqTileChunks chunks (coalesced) into shared memory, zero padded past n.
__device__ __forceinline__ void seq_stage(const double* w, long long n, long long first_elem, double* tile)
{
    for (int i = threadIdx.x; i < kSeqTileChunks * kL1; i += blockDim.x) {
        const long long g = first_elem + i;
        tile[(i / kL1) * kSeqRowPitch + (i % kL1)] = g < n ? w[g] : 0.0;
    }
}

// S1: plain per-chunk sums and per-tile totals (approximate: they only steer the binade guess).
__global__ void __launch_bounds__(128) seq_chunk_sums_kernel(const double* w, long long n, long long n1, double* sums,
                                                             double* tile_sums)
{
    __shared__ double tile[kSeqTileChunks * kSeqRowPitch];
    const long long chunk0 = (long long)blockIdx.x * kSeqTileChunks;
    seq_stage(w, n, chunk0 * kL1, tile);
    __syncthreads();
    if (threadIdx.x < kSeqTileChunks) {                 // warp 0: one lane per chunk
        double s = 0.0;
        if (chunk0 + threadIdx.x < n1) {
            const double* row = tile + threadIdx.x * kSeqRowPitch;
#pragma unroll 8
            for (int i = 0; i < kL1; ++i) s += row[i];
            sums[chunk0 + threadIdx.x] = s;
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
    }
}

// S2: exclusive scan of the per-tile totals (single CTA; a tile is 32 chunks, so this is N/4096 values).
__global__ void __launch_bounds__(1024) seq_tile_scan_kernel(const double* tile_sums, long long ntiles, double* tile_excl)
{
    __shared__ double warp_tot[32];
    __shared__ double carry_s;
    if (threadIdx.x == 0) carry_s = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long base = 0; base < ntiles; base += 1024) {
        const long long k = base + threadIdx.x;
        const double v = k < ntiles ? tile_sums[k] : 0.0;
        double inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            double t = warp_tot[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const double u = __shfl_up_sync(0xffffffffu, t, off);
                if (lane >= off) t += u;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const double incl = carry_s + (wid ? warp_tot[wid - 1] : 0.0) + inc;
        if (k < ntiles) tile_excl[k] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

// S3: per level-1 chunk, Q[p] = (c_out - c_in)/ulp for entry parity p, from two real double-add chains started at the
// binade's two representatives.  A chain that leaves the binade invalidates the chunk (ebias := 0).
// The binade guess comes first: ebias[k] = biased exponent shared by the chunk's whole running sum according to the
// approximate prefix (tile prefix + warp scan of the tile's chunk sums, with a 1e-6 relative safety margin), or 0 =
// "walk it serially".
__global__ void __launch_bounds__(128) seq_chunk_maps_kernel(const double* w, long long n, long long n1,
                                                             const double* sums, const double* tile_excl, int* ebias,
                                                             long long* q0, long long* q1)
{
    __shared__ double tile[kSeqTileChunks * kSeqRowPitch];
    const long long chunk0 = (long long)blockIdx.x * kSeqTileChunks;
    seq_stage(w, n, chunk0 * kL1, tile);
    __syncthreads();
    const long long k = chunk0 + threadIdx.x;
    if (threadIdx.x < kSeqTileChunks) {                 // warp 0, all 32 lanes (shuffles below)
        const double v = k < n1 ? sums[k] : 0.0;
        double inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, off);
            if ((int)threadIdx.x >= off) inc += t;
        }
        const double incl = tile_excl[blockIdx.x] + inc, excl = incl - v;
        const double lo = excl * (1.0 - 1e-6), hi = incl * (1.0 + 1e-6);
        const int elo = dbl_exp(lo), ehi = dbl_exp(hi);
        const int e = (lo > 0.0 && elo == ehi && elo > 60 && elo < 1900) ? elo : 0;
        if (k < n1) ebias[k] = e;
        if (k < n1 && e != 0) {
            const double base0 = __longlong_as_double((long long)e << 52);
            const double ulp = __longlong_as_double((long long)(e - 52) << 52);
            const double base1 = base0 + ulp;
            const double top = base0 + base0;
            const double* row = tile + threadIdx.x * kSeqRowPitch;
            double c0 = base0, c1 = base1;
#pragma unroll 8
            for (int i = 0; i < kL1; ++i) {
                const double v = row[i];
                c0 = __dadd_rn(c0, v);
                c1 = __dadd_rn(c1, v);
            }
            if (c0 < top && c1 < top && c0 >= base0 && c1 >= base1) {
                // exact: both differences are multiples of ulp below 2^52 ulp
                q0[k] = __double_as_longlong(c0) - __double_as_longlong(base0);   // same binade: bit patterns count ulps
                q1[k] = __double_as_longlong(c1) - __double_as_longlong(base1);
            } else {
                ebias[k] = 0;
            }
        }
    }
}

// S4: compose the level-1 maps of each level-2 group (serial over kL2 entries per thread: tiny).
// gebias = common exponent or 0; (g0, g1) = composed map.
__global__ void seq_group_maps_kernel(const int* ebias, const long long* q0, const long long* q1, long long n1,
                                      long long n2, int* gebias, long long* g0, long long* g1)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n2) return;
    const long long first = j * kL2;
    const long long last = first + kL2 < n1 ? first + kL2 : n1;
    const int e = ebias[first];
    bool ok = e != 0;
    long long a0 = 0, a1 = 0;     // running composed map for entry parity 0 / 1
    for (long long k = first; k < last && ok; ++k) {
        if (ebias[k] != e) { ok = false; break; }
        const long long f0 = q0[k], f1 = q1[k];
        a0 += ((a0 & 1) ? f1 : f0);
        a1 += (((a1 + 1) & 1) ? f1 : f0);   // entry parity 1: current parity = (1 + a1) & 1
    }
    gebias[j] = ok ? e : 0;
    g0[j] = a0;
    g1[j] = a1;
}

// Apply a verified map to an exact running sum: returns false if the entry binade differs or the sum would leave it.
__device__ __forceinline__ bool seq_apply(double& c, int e, long long m0, long long m1)
{
    if (e == 0 || dbl_exp(c) != e) return false;
    const long long bits = __double_as_longlong(c);
    const long long add = (bits & 1) ? m1 : m0;
    const long long nb = bits + add;                // same binade <=> exponent field unchanged
    if (((nb >> 52) & 0x7ff) != e) return false;
    c = __longlong_as_double(nb);
    return true;
}

// Composition of two "add D[parity]" maps: first a, then b.
__device__ __forceinline__ void seq_compose(long long a0, long long a1, long long b0, long long b1, long long& r0,
                                            long long& r1)
{
    r0 = a0 + ((a0 & 1) ? b1 : b0);
    r1 = a1 + (((a1 + 1) & 1) ? b1 : b0);
}

// Tries to advance the exact running sum c over 32 consecutive maps at once (lane l holds map l: exponent e_l and
// (m0_l, m1_l); lanes >= cnt hold identity maps).  Succeeds iff every map assumes c's binade and the sum stays inside it
// through the last one; then cin_l = exact running sum at the entry of map l and c = sum after the last map.
__device__ __forceinline__ bool seq_apply_warp(double& c, int cnt, int e_l, long long m0_l, long long m1_l, double& cin_l)
{
    const int lane = threadIdx.x & 31;
    const int e = dbl_exp(c);
    const bool mine_ok = lane >= cnt || (e_l == e && e_l != 0);
    if (!__all_sync(0xffffffffu, mine_ok)) return false;
    long long p0 = lane < cnt ? m0_l : 0, p1 = lane < cnt ? m1_l : 0;      // inclusive prefix composition
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const long long q0 = __shfl_up_sync(0xffffffffu, p0, off), q1 = __shfl_up_sync(0xffffffffu, p1, off);
        if (lane >= off) {
            long long r0, r1;
            seq_compose(q0, q1, p0, p1, r0, r1);
            p0 = r0; p1 = r1;
        }
    }
    const long long bits = __double_as_longlong(c);
    const long long incl = bits + ((bits & 1) ? p1 : p0);
    long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = bits;
    // monotone: if the last sum is still in the binade, all are
    const long long last = __shfl_sync(0xffffffffu, incl, 31);
    if (((last >> 52) & 0x7ff) != e || last < bits) return false;
    cin_l = __longlong_as_double(excl);
    c = __longlong_as_double(last);
    return true;
}

// S5: the walk over the level-2 groups (one warp).  Runs of groups that stay inside one binade advance 32 at a time
// (seq_apply_warp); the rest one by one, opening a group into its level-1 chunks where its map does not apply and
// adding raw elements where a chunk's does not.  All lanes carry the same c, so shuffles broadcast staged values.
// Produces the exact running sum at the entry of every level-2 group (cin2), of every level-1 chunk of groups that
// had to be opened (cin1, flagged in opened[j]), the exact total, and the number of chunks that took raw adds.
__global__ void __launch_bounds__(1024) seq_walk_kernel(const double* w, long long n, long long n1, long long n2,
                                                      const int* ebias, const long long* q0, const long long* q1,
                                                      const int* gebias, const long long* g0, const long long* g1,
                                                      double* cin2, double* cin1, int* opened, double* total,
                                                      long long* fallback_chunks, int staged)
{
    // The walk itself is one warp's work, but its loads sit on the critical path: the whole CTA first stages the group
    // maps in shared memory (staged = 1: they fit), so each step of the walk costs shared-memory, not DRAM, latency.
    __shared__ double fb_chunk[kL1];          // the raw elements of a chunk that has to be added one by one
    extern __shared__ __align__(16) unsigned char walk_smem[];
    long long* sg0 = reinterpret_cast<long long*>(walk_smem);
    long long* sg1 = sg0 + (staged ? n2 : 0);
    int* sge = reinterpret_cast<int*>(sg1 + (staged ? n2 : 0));
    if (staged) {
        for (long long j = threadIdx.x; j < n2; j += blockDim.x) { sg0[j] = g0[j]; sg1[j] = g1[j]; sge[j] = gebias[j]; }
        __syncthreads();
    }
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    double c = 0.0;
    long long fallbacks = 0;
    for (long long jb = 0; jb < n2; jb += 32) {
        const long long jl = jb + lane;
        const int ge_l = jl < n2 ? (staged ? sge[jl] : gebias[jl]) : 0;
        const long long g0_l = jl < n2 ? (staged ? sg0[jl] : g0[jl]) : 0, g1_l = jl < n2 ? (staged ? sg1[jl] : g1[jl]) : 0;
        const int cnt = (int)(n2 - jb < 32 ? n2 - jb : 32);
        double cin_l;
        if (seq_apply_warp(c, cnt, ge_l, g0_l, g1_l, cin_l)) {
            if (lane < cnt) { cin2[jl] = cin_l; opened[jl] = 0; }
            continue;
        }
        for (int t = 0; t < cnt; ++t) {
            const long long j = jb + t;
            const int ge = __shfl_sync(0xffffffffu, ge_l, t);
            const long long m0 = __shfl_sync(0xffffffffu, g0_l, t), m1 = __shfl_sync(0xffffffffu, g1_l, t);
            if (lane == 0) cin2[j] = c;
            if (seq_apply(c, ge, m0, m1)) {
                if (lane == 0) opened[j] = 0;
                continue;
            }
            if (lane == 0) opened[j] = 1;
            const long long kfirst = j * kL2;
            const long long klast = kfirst + kL2 < n1 ? kfirst + kL2 : n1;
            for (long long kb = kfirst; kb < klast; kb += 32) {
                const long long kl = kb + lane;
                const int e_l = kl < klast ? ebias[kl] : 0;
                const long long f0_l = kl < klast ? q0[kl] : 0, f1_l = kl < klast ? q1[kl] : 0;
                const int kc = (int)(klast - kb < 32 ? klast - kb : 32);
                if (seq_apply_warp(c, kc, e_l, f0_l, f1_l, cin_l)) {
                    if (lane < kc) cin1[kl] = cin_l;
                    continue;
                }
                for (int s = 0; s < kc; ++s) {
                    const long long k = kb + s;
                    const int e = __shfl_sync(0xffffffffu, e_l, s);
                    const long long f0 = __shfl_sync(0xffffffffu, f0_l, s), f1 = __shfl_sync(0xffffffffu, f1_l, s);
                    if (lane == 0) cin1[k] = c;
                    if (seq_apply(c, e, f0, f1)) continue;
                    ++fallbacks;
                    const long long efirst = k * kL1;
                    double v_l[kL1 / 32];
#pragma unroll
                    for (int q = 0; q < kL1 / 32; ++q) {                      // all loads in flight before the adds
                        const long long gi = efirst + q * 32 + lane;
                        v_l[q] = gi < n ? w[gi] : 0.0;
                    }
                    // the adds are one dependent chain either way; reading the staged values back from shared memory
                    // (a broadcast, every lane keeps the same c) keeps the chain at DADD latency instead of
                    // shuffle + DADD latency per element
#pragma unroll
                    for (int q = 0; q < kL1 / 32; ++q) fb_chunk[q * 32 + lane] = v_l[q];
                    __syncwarp();
#pragma unroll 16
                    for (int i = 0; i < kL1; ++i) c = __dadd_rn(c, fb_chunk[i]);
                    __syncwarp();
                }
            }
        }
    }
    if (lane == 0) {
        *total = c;
        *fallback_chunks = fallbacks;
    }
}

// S6: entry sums of the level-1 chunks inside groups the walk did not open (all in one binade, verified by S5).
__global__ void seq_group_expand_kernel(const long long* q0, const long long* q1, long long n1, long long n2,
                                        const double* cin2, const int* opened, double* cin1)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n2 || opened[j]) return;
    const long long first = j * kL2;
    const long long last = first + kL2 < n1 ? first + kL2 : n1;
    long long bits = __double_as_longlong(cin2[j]);
    for (long long k = first; k < last; ++k) {
        cin1[k] = __longlong_as_double(bits);
        bits += (bits & 1) ? q1[k] : q0[k];
    }
}

// S7: materialise the exact running sum c_i for every element: real sequential double adds inside each chunk.
__global__ void __launch_bounds__(128) seq_materialize_kernel(const double* w, long long n, long long n1,
                                                              const double* cin1, double* cum)
{
    __shared__ double tile[kSeqTileChunks * kSeqRowPitch];
    const long long chunk0 = (long long)blockIdx.x * kSeqTileChunks;
    seq_stage(w, n, chunk0 * kL1, tile);
    __syncthreads();
    const long long k = chunk0 + threadIdx.x;
    if (threadIdx.x < kSeqTileChunks && k < n1) {
        double* row = tile + threadIdx.x * kSeqRowPitch;
        double c = cin1[k];
#pragma unroll 8
        for (int i = 0; i < kL1; ++i) {
            c = __dadd_rn(c, row[i]);
            row[i] = c;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSeqTileChunks * kL1; i += blockDim.x) {
        const long long g = chunk0 * kL1 + i;
        if (g < n) cum[g] = tile[(i / kL1) * kSeqRowPitch + (i % kL1)];
    }
}

// Systematic search: child m takes the first i with !(U_m > c_i), U_m = r + m*(1/N) evaluated exactly like the
// reference (double product, double sum; particle_filter.cpp:89,95-99).  c is non-decreasing, so the reference's
// forward-only loop equals a lower-bound search.  Past-the-end (the reference's unbounded loop) clamps to N-1.
constexpr int kSearchRun = 8;       // consecutive children per thread
__global__ void resample_search_kernel(const double* cum, long long n, double r, long long lo, long long hi,
                                       int32_t* idx, unsigned long long* overruns)
{
    // A thread takes kSearchRun consecutive children: one lower-bound search for the first, then the reference's own
    // forward walk (U grows by 1/N per child and c is non-decreasing, so the next index is the same or a few further);
    // a walk longer than 32 steps falls back to a search.  Same indices as a search per child, a sixth of the loads.
    const double m_inv = __ddiv_rn(1.0, (double)n);
    const long long runs = (hi - lo + kSearchRun - 1) / kSearchRun;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < runs; t += (long long)gridDim.x * blockDim.x) {
        const long long m0 = lo + t * kSearchRun;
        const long long m1 = m0 + kSearchRun < hi ? m0 + kSearchRun : hi;
        long long a = 0;
        for (long long m = m0; m < m1; ++m) {
            const double u = __dadd_rn(r, __dmul_rn((double)m, m_inv));
            int steps = 0;
            if (m != m0)
                while (a < n && u > cum[a] && steps < 32) { ++a; ++steps; }
            if (m == m0 || steps == 32) {
                long long b = n;                  // lower bound in [a, n): c is non-decreasing and so is u
                while (a < b) {
                    const long long mid = (a + b) >> 1;
                    if (u > cum[mid]) a = mid + 1; else b = mid;
                }
            }
            long long out = a;
            if (out >= n) { out = n - 1; atomicAdd(overruns, 1ull); }
            idx[m] = (int32_t)out;
        }
    }
}

// =================================================================================================================
// The fast pass's view of the map.  The score only ever distinguishes "cell > 0" (its value matters) from "cell <= 0"
// (it reads as nothing), so the non-positive cells of this derived copy are free to carry one bit of look-ahead:
//   -1 : the cell and its whole 5 x 5 neighbourhood are non-positive.  Wherever the float pass lands within one cell of
//        the reference's endpoint (its error is far below a cell), the reference's endpoint cell and both Bresenham
//        neighbours lie inside that neighbourhood, so the ray's score is 0 whatever the exact cell and octant are --
//        certain without the boundary and direction tests;
//    0 : non-positive, but something positive is within two cells;
//   >0 : the cell's own value.
// Recomputed for the rectangle (+2 cells) of every map change (mcl_set_map, mcl_update_map_rect, mcl_map_update).
// =================================================================================================================
__global__ void derive_fast_map_kernel(const int8_t* cells, int8_t* out, int width, int height, int pitch, int x0,
                                       int y0, int w, int h)
{
    const int total = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = y0 + i / w, x = x0 + i % w;
        const int v = cells[(size_t)y * pitch + x];
        int r = v;
        if (v <= 0) {
            bool any = false;
            for (int dy = -2; dy <= 2; ++dy) {
                const int yy = y + dy;
                if ((unsigned)yy >= (unsigned)height) continue;
                for (int dx = -2; dx <= 2; ++dx) {
                    const int xx = x + dx;
                    if ((unsigned)xx < (unsigned)width) any = any || cells[(size_t)yy * pitch + xx] > 0;
                }
            }
            r = any ? 0 : -1;
        }
        out[(size_t)y * pitch + x] = (int8_t)r;
    }
}

// =================================================================================================================
// Obstacle-distance grid of the mirror: planning/obstacle_distance_grid.cpp:73-188 runs a four-connected brushfire from
// every cell with log-odds >= 0 (occupied OR unknown) over the free cells, all steps costing the same, so a free cell's
// distance is its Manhattan distance (in steps) to the nearest such cell -- which two separable sweeps compute exactly:
// down and up every column, then left and right along every row.  thr selects the sources: cells >= thr (0 = the
// reference's rule; 1 = occupied cells only, for the likelihood field).  65535 = no source anywhere.
// =================================================================================================================
constexpr unsigned kDtInf = 65535u;
__global__ void dt_columns_kernel(const int8_t* cells, int width, int height, int pitch, int thr, uint16_t* steps)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    unsigned d = kDtInf;
    for (int y = 0; y < height; ++y) {
        d = (int)cells[(size_t)y * pitch + x] >= thr ? 0u : min(d + 1u, kDtInf);
        steps[(size_t)y * width + x] = (uint16_t)d;
    }
    d = kDtInf;
    for (int y = height - 1; y >= 0; --y) {
        d = min((unsigned)steps[(size_t)y * width + x], min(d + 1u, kDtInf));
        steps[(size_t)y * width + x] = (uint16_t)d;
    }
}
__global__ void dt_rows_kernel(int width, int height, uint16_t* steps)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= height) return;
    uint16_t* row = steps + (size_t)y * width;
    unsigned d = kDtInf;
    for (int x = 0; x < width; ++x) { d = min((unsigned)row[x], min(d + 1u, kDtInf)); row[x] = (uint16_t)d; }
    d = kDtInf;
    for (int x = width - 1; x >= 0; --x) { d = min((unsigned)row[x], min(d + 1u, kDtInf)); row[x] = (uint16_t)d; }
}
// steps -> the reference's float distances: d_k = fl(d_{k-1} + 0.1f) (obstacle_distance_grid.cpp:179; the table is
// built on the host), -1 where the brushfire never arrives.
__global__ void dt_to_float_kernel(const uint16_t* steps, long long count, const float* table, int table_len, float* out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const unsigned d = steps[i];
        out[i] = d == kDtInf ? -1.0f : table[min((int)d, table_len - 1)];
    }
}
// Likelihood field of the non-parity sensor mode: u = max(0, 127 - 8 d^2), d = steps to the nearest OCCUPIED cell
// (> 0): 127 on a wall, 119 / 95 / 55 one / two / three steps away, 0 beyond.
__host__ __device__ inline int lf_value(unsigned d) { return d >= 4u ? 0 : 127 - 8 * (int)(d * d); }
__global__ void lf_field_kernel(const uint16_t* steps, int width, int height, int pitch, int8_t* lf)
{
    const long long total = (long long)width * height;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / width), x = (int)(i - (long long)y * width);
        lf[(size_t)y * pitch + x] = (int8_t)lf_value(steps[i]);
    }
}

// =================================================================================================================
// K5  map update on the device mirror: Mapping::updateMap (mapping.cpp:17-127).
// The reference raises the endpoint cell of every ray by hitOdds (saturating at 127), THEN lowers every cell of each
// ray's Bresenham walk (start cell up to, not including, the endpoint cell) by missOdds (saturating at -128), one ray
// after the other.  Saturating adds of one sign commute, so the parallel form is exact: count, per cell, how many
// endpoints (high half-word) and how many walk visits (low half-word) it receives, then apply
//     v = max(-128, min(127, v + hits*hitOdds) - visits*missOdds).
// The rays are the MovingLaserScan between the previous and the current SLAM pose, built with the same
// exactly-rounded arithmetic as the sensor model (exact_endpoint).
// =================================================================================================================
struct MapUpdateArgs {
    const Beam* beams;          // valid beams (range > min_range) with ratios between the two poses' utimes
    int num_beams;
    float xa, ya, tha, xb, yb, thb;   // current pose (end of sweep) / previous pose
    float max_laser_distance;
    DevGrid grid;
    int wx0, wy0, ww, wh;       // count window (grid cells); every touched in-grid cell must fall inside
    uint32_t* counts;           // [wh][ww], zeroed
    int* error_flag;            // set when a ray leaves the count window or has unusable coordinates
};

template <bool INTERP>
__global__ void __launch_bounds__(128) map_count_kernel(const MapUpdateArgs a)
{
    GridConst gc;
    gc.gx = (double)a.grid.origin_x; gc.gy = (double)a.grid.origin_y;
    gc.cpm = a.grid.cells_per_meter; gc.cpm_d = (double)a.grid.cells_per_meter;
    gc.trig = gs_load_consts();
    const RayBase rb = make_ray_base(a.xa, a.ya, a.tha, a.xb, a.yb, a.thb);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.num_beams; j += gridDim.x * blockDim.x) {
        const Beam b = a.beams[j];
        if (!(b.range <= a.max_laser_distance)) continue;                    // mapping.cpp:43, :60
        float sx, sy, px, py, e1x, e1y;
        exact_endpoint<INTERP>(rb, b, gc, sx, sy, px, py, e1x, e1y);
        const int cx = f2i_x86(e1x), cy = f2i_x86(e1y);                       // :48-49
        int x = f2i_x86(sx), y = f2i_x86(sy);                                 // bresenham(rayStart.x, ...) :68
        // the reference would walk (practically) forever on such coordinates; report instead
        if (abs(cx) > (1 << 20) || abs(cy) > (1 << 20) || abs(x) > (1 << 20) || abs(y) > (1 << 20)) {
            atomicExch(a.error_flag, 1);
            continue;
        }
        auto bump = [&](int gx, int gy, uint32_t inc) {
            if ((unsigned)gx >= (unsigned)a.grid.width || (unsigned)gy >= (unsigned)a.grid.height) return;   // isCellInGrid
            const int tx = gx - a.wx0, ty = gy - a.wy0;
            if ((unsigned)tx >= (unsigned)a.ww || (unsigned)ty >= (unsigned)a.wh) { atomicExch(a.error_flag, 2); return; }
            atomicAdd(a.counts + (size_t)ty * a.ww + tx, inc);
        };
        bump(cx, cy, 1u << 16);                                               // scoreEndpoint :42-57
        const int dx = abs(cx - x), dy = abs(cy - y);                         // bresenham :101-127
        const int stepx = x < cx ? 1 : -1, stepy = y < cy ? 1 : -1;
        int err = dx - dy;
        while (x != cx || y != cy) {
            bump(x, y, 1u);
            const int e2 = 2 * err;        // the reference compares (float)(2*err): exact below 2^24
            if (e2 >= -dy) { err -= dy; x += stepx; }
            if (e2 <= dx) { err += dx; y += stepy; }
        }
    }
}

__global__ void map_apply_kernel(int8_t* cells, int pitch, int wx0, int wy0, int ww, int wh, const uint32_t* counts,
                                 int hit, int miss)
{
    const int total = ww * wh;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t c = counts[i];
        if (c == 0) continue;
        const int ty = i / ww, tx = i - ty * ww;
        int8_t* cell = cells + (size_t)(wy0 + ty) * pitch + (wx0 + tx);
        int v = *cell;
        v = min(127, v + (int)(c >> 16) * hit);            // increaseCellOdds x hits   (mapping.cpp:73-85)
        v = max(-128, v - (int)(c & 0xffffu) * miss);      // decreaseCellOdds x visits (:87-99)
        *cell = (int8_t)v;
    }
}

// =================================================================================================================
// Initialisers and utilities.
// =================================================================================================================
// particle_filter.cpp:16-34 with Philox normals: x,y,theta ~ pose + N(0, std); last particle = exact pose (:33).
__global__ void init_at_pose_kernel(float* x, float* y, float* th, float* px, float* py, float* pth, long long n,
                                    float x0, float y0, float th0, double std, uint64_t seed)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), 0u, 0x494e4954u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        float z0, z1, z2, z3;
        philox_normals(w, z0, z1, z2, z3);
        const float xs = (float)__dadd_rn((double)x0, __dmul_rn((double)z0, std));
        const float ys = (float)__dadd_rn((double)y0, __dmul_rn((double)z1, std));
        const float ts = wrap_to_pi((float)__dadd_rn((double)th0, __dmul_rn((double)z2, std)));
        px[i] = xs; py[i] = ys; pth[i] = ts;
        const bool last = i == n - 1;
        x[i] = last ? x0 : xs; y[i] = last ? y0 : ys; th[i] = last ? th0 : ts;
    }
}

__global__ void init_uniform_kernel(float* x, float* y, float* th, float* px, float* py, float* pth, long long n,
                                    float gx, float gy, float wm, float hm, uint64_t seed, int nbx, int nby)
{
    // stratified: particle i belongs to block b = floor(i * NB / n); blocks are visited row by row, alternate rows
    // right-to-left, so consecutive blocks (and therefore consecutive particles) are always neighbours
    const long long nblk = (long long)nbx * nby;
    const float bwm = wm / (float)nbx, bhm = hm / (float)nby;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), 0u, 0x554e4946u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const long long b = (long long)(((unsigned __int128)(unsigned long long)i * (unsigned long long)nblk) / (unsigned long long)n);
        const int by = (int)(b / nbx);
        int bx = (int)(b - (long long)by * nbx);
        if (by & 1) bx = nbx - 1 - bx;
        const float k = 2.3283064365386963e-10f;
        const float xs = gx + bwm * ((float)bx + (float)w.x * k);
        const float ys = gy + bhm * ((float)by + (float)w.y * k);
        const float ts = wrap_to_pi((float)(((double)w.z * 2.3283064365386963e-10 - 0.5) * kTwoPi));
        x[i] = px[i] = xs; y[i] = py[i] = ys; th[i] = pth[i] = ts;
    }
}

// AoS particle_t <-> SoA
struct AosParticle { long long utime; float x, y, th; int pad0; long long putime; float px, py, pth; int pad1; double w; };
static_assert(sizeof(AosParticle) == 56, "particle_t layout");

__global__ void aos_to_soa_kernel(const AosParticle* aos, long long n, float* x, float* y, float* th, float* px,
                                  float* py, float* pth, double* w)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const AosParticle p = aos[i];
        x[i] = p.x; y[i] = p.y; th[i] = p.th; px[i] = p.px; py[i] = p.py; pth[i] = p.pth; w[i] = p.w;
    }
}

__global__ void soa_to_aos_kernel(AosParticle* aos, long long count, long long stride, long long utime,
                                  long long putime, const float* x, const float* y, const float* th, const float* px,
                                  const float* py, const float* pth, const double* w)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < count; k += (long long)gridDim.x * blockDim.x) {
        const long long i = k * stride;
        AosParticle p;
        p.utime = utime; p.x = x[i]; p.y = y[i]; p.th = th[i]; p.pad0 = 0;     // padding bytes exported as zeros
        p.putime = putime; p.px = px[i]; p.py = py[i]; p.pth = pth[i]; p.pad1 = 0;
        p.w = w[i];
        aos[k] = p;
    }
}

// particles_t export of a drawn sub-sample: particle k = cloud[idx[k]] with weight w_out (a weighted draw carries
// equal weights).
__global__ void soa_to_aos_idx_kernel(AosParticle* aos, long long count, const int32_t* idx, double w_out, long long utime,
                                      long long putime, const float* x, const float* y, const float* th, const float* px,
                                      const float* py, const float* pth)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < count; k += (long long)gridDim.x * blockDim.x) {
        const long long i = idx[k];
        AosParticle p;
        p.utime = utime; p.x = x[i]; p.y = y[i]; p.th = th[i]; p.pad0 = 0;
        p.putime = putime; p.px = px[i]; p.py = py[i]; p.pth = pth[i]; p.pad1 = 0;
        p.w = w_out;
        aos[k] = p;
    }
}

__global__ void score_to_double_kernel(const int32_t* score2, double* out, long long n)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = 0.5 * (double)score2[i];
}

// Bounding box of the poses in [lo, hi) as ordered-int atomics: box = (minx, miny, maxx, maxy) of float bit patterns
// mapped to a monotone integer order.
__device__ __forceinline__ int float_order(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__global__ void bbox_kernel(const float* x, const float* y, const float* px, const float* py, long long lo,
                            long long hi, int* box)
{
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
    for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi;
         i += (long long)gridDim.x * blockDim.x) {
        const int a = float_order(x[i]), b = float_order(y[i]), c = float_order(px[i]), d = float_order(py[i]);
        mnx = min(mnx, min(a, c)); mxx = max(mxx, max(a, c));
        mny = min(mny, min(b, d)); mxy = max(mxy, max(b, d));
    }
    for (int off = 16; off > 0; off >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, off));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, off));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, off));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(box + 0, mnx); atomicMin(box + 1, mny);
        atomicMax(box + 2, mxx); atomicMax(box + 3, mxy);
    }
}

// Roofline denominator: uniformly random 1-byte reads, L1 bypassed (ld.global.cg), one per lane per step.
__global__ void gather_peak_kernel(const int8_t* buf, unsigned long long mask, long long reads_per_thread,
                                   unsigned long long* sink)
{
    unsigned long long s = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 1;
    int acc = 0;
    for (long long i = 0; i < reads_per_thread; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;          // xorshift64
        acc += (int)__ldcg(buf + (s & mask));
    }
    if (acc == 0x7fffffff) *sink = (unsigned long long)acc;
}

// Certification margin probe: over every (particle, beam) the fast pass would evaluate, the largest deviation between
// the float model and the reference's exactly-rounded values, for the endpoint (dev[0]) and for the extended point
// (dev[1]); the host compares them with the eps the certification assumes.  Floats >= 0 order like their bit patterns.
template <bool INTERP>
__global__ void fast_margin_kernel(const ScoreArgs a, unsigned* dev_bits)
{
    extern __shared__ __align__(16) unsigned char smem[];
    Beam* sbeams = reinterpret_cast<Beam*>(smem);
    for (int i = threadIdx.x; i < a.num_beams; i += blockDim.x) sbeams[i] = a.beams[i];
    __syncthreads();
    GridConst gc;
    gc.gx = (double)a.grid.origin_x; gc.gy = (double)a.grid.origin_y;
    gc.cpm = a.grid.cells_per_meter; gc.cpm_d = (double)a.grid.cells_per_meter;
    gc.trig = gs_load_consts();
    float de = 0.0f, d2 = 0.0f;
    for (long long p = a.lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; p < a.hi;
         p += (long long)gridDim.x * blockDim.x) {
        const RayBase rb = make_ray_base(a.x[p], a.y[p], a.th[p], a.px[p], a.py[p], a.pth[p]);
        const FastBase fb = make_fast_base<INTERP>(a.x[p], a.y[p], a.th[p], a.px[p], a.py[p], a.pth[p], gc.gx, gc.gy,
                                                   gc.cpm_d, a.fast);
        if (!fb.ok) continue;
        for (int j = 0; j < a.num_beams; ++j) {
            const Beam b = sbeams[j];
            FastBeam f;
            f.ratio = (float)b.ratio; f.theta = b.theta; f.rc = __fmul_rn(b.range, gc.cpm); f.pad = 0.0f;
            float fpx, fpy, fex, fey, sx, sy, px, py, e1x, e1y;
            fast_endpoint<INTERP>(fb, f, fpx, fpy, fex, fey);
            exact_endpoint<INTERP>(rb, b, gc, sx, sy, px, py, e1x, e1y);
            // the budget covers endpoints the fast pass can certify: inside the window's certain-interior box, or
            // beyond the grid (there only the side of the boundary matters, checked with the same eps)
            if (!(fabsf(fex) <= 8192.0f && fabsf(fey) <= 8192.0f)) continue;
            de = fmaxf(de, fmaxf(fabsf(fex - e1x), fabsf(fey - e1y)));
            const float x2 = __fadd_rn(__fmul_rn(2.0f, px), sx), y2 = __fadd_rn(__fmul_rn(2.0f, py), sy);
            d2 = fmaxf(d2, fmaxf(fabsf(__fadd_rn(fex, fpx) - x2), fabsf(__fadd_rn(fey, fpy) - y2)));
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        de = fmaxf(de, __shfl_xor_sync(0xffffffffu, de, off));
        d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(dev_bits + 0, __float_as_uint(de));
        atomicMax(dev_bits + 1, __float_as_uint(d2));
    }
}

// Largest |fast_sincos - sin/cos in double| over every float in [lo, hi] (walks the bit patterns: both signs).
__global__ void fast_trig_error_kernel(float lo, float hi, unsigned long long* out_bits /* [2]: max errs as double bits */)
{
    double es = 0.0, ec = 0.0;
    // non-negative floats up to max(|lo|, |hi|), each tried with both signs where inside [lo, hi]
    const unsigned top = __float_as_uint(fmaxf(fabsf(lo), fabsf(hi)));
    for (unsigned long long b = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; b <= top;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float v = __uint_as_float((unsigned)b);
#pragma unroll
        for (int sgn = 0; sgn < 2; ++sgn) {
            const float a = sgn ? -v : v;
            if (a < lo || a > hi) continue;
            float s, c;
            fast_sincos(a, &s, &c);
            double sd, cd;
            sincos((double)a, &sd, &cd);
            es = fmax(es, fabs((double)s - sd));
            ec = fmax(ec, fabs((double)c - cd));
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        es = fmax(es, __shfl_xor_sync(0xffffffffu, es, off));
        ec = fmax(ec, __shfl_xor_sync(0xffffffffu, ec, off));
    }
    if ((threadIdx.x & 31) == 0) {      // non-negative doubles order like their bit patterns
        atomicMax(out_bits + 0, (unsigned long long)__double_as_longlong(es));
        atomicMax(out_bits + 1, (unsigned long long)__double_as_longlong(ec));
    }
}

__global__ void debug_sincosf_kernel(const float* x, long long n, float* s, float* c)
{
    // the exact variant the sensor kernel runs, plus the signed-zero fix-up of glibc_sincosf_fast
    const GsConsts k = gs_load_consts();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float sv, cv;
        glibc_sincosf_regs(k, x[i], &sv, &cv);
        s[i] = (x[i] == 0.0f) ? x[i] : sv;
        c[i] = cv;
    }
}

}  // namespace mcl
