// Single-pass sensor model with a per-window SCORE TABLE in shared memory (sm_100a).
//
// sensor_model.cpp:28-59 turns a ray into at most three cell reads selected by (a) the endpoint cell and (b) the octant
// class of the ray direction (the two one-step Bresenham moves of :61-86).  Both selections only depend on the map, so
// the CTA resolves them ONCE per window instead of once per evaluation:
//
//   K[cell]  (uint16 tile)   0          nothing positive within two cells (5x5 empty): the ray scores 0 whatever the
//                                       exact cell and octant are
//                            1          non-positive cell with no positive 8-neighbour: scores 0 once the cell is certain
//                            2..128     occupied cell, log-odds K-1: scores 2*(K-1) half units once the cell is certain
//                            >= 129     non-positive cell with a positive neighbour: entry K of T
//   T[K][s]  (8 bytes/entry) the ray's score in half units when the endpoint cell is that cell and the direction lies
//                            in sector s (0: +x, 1: +x+y, 2: +y, ... 7: +x-y): o1 > 0 ? o1 : max(o2, 0) with
//                            o1 = cell one step against s (toward the robot), o2 = cell one step along s (toward the
//                            doubled endpoint).  Entries 0..128 are the uniform classes (same byte in all 8 slots),
//                            so one dependent LDS gives the score of every class.
//
// One evaluation is then: the float model of the endpoint in WINDOW-NORMALISED coordinates ((cell + kappa - 0.5) /
// (w - 1.5), so one saturating FFMA both evaluates it and clamps it onto the window -- rays that leave a clipped window
// land on its border cells, which are class 0 -- and NaNs become 0), +1.0 to make the mantissa linear, one IMAD.WIDE
// per axis by (2w - 3) 2^8 that yields the cell in the high word and the 32-bit fraction in the low word, K, the sector
// from the angle by magic-number rounding (boundaries at +-atan(1/2) around each axis), T -- about half the issue
// slots of score_beam_fast, with the compare/select work (the ALU pipe, the co-limiter there) cut to a third.
//
// Certification (the score is only taken when it provably equals the reference's):
//   * cell:   the coordinate is further than kappa from an integer, kappa >= eps + the float roundings of the
//             normalised coordinates (TabPlan / table_plan_kernel has the budget);
//   * sector: the angle is further than d8 (per beam: asin(T3 / (sqrt5 rc)) + slack) from the octant boundaries, i.e.
//             |2|px| - |py|| and |2|py| - |px|| exceed T3 = 3(1 + 2 eps) -- only needed for K >= 129;
//   * K == 0 needs neither (mcl_kernels.cuh: derive_fast_map_kernel has the argument).
// Everything else is DEFERRED to the literal restatement (score_beam) -- in the same kernel: every lane keeps one bit per
// beam of the current 32-beam word, the warp compacts the set bits of its 32 lanes into a shared-memory queue and drains
// it 32 entries at a time (poses travel by shuffle, results by shared-memory integer atomics), so there is no mask
// array in HBM and no second launch.
//
// The window is the bounding box of the cloud +- the longest ray, clipped to the grid plus four cells: cells outside the
// grid read 0 (occupancy_grid.cpp:65-70), so an endpoint beyond the grid's high edges is just another table lookup, and
// four or more cells out everything is class 0.  Below zero the reference truncates toward zero instead of flooring (sensor_model.cpp:34-38 casts to int), so window cells
// with a negative global coordinate are never certified (class 1 + a cell test) unless they are four or more cells out
// (class 0: the truncated cell and its neighbours are all outside the grid).
//
// The window is planned ON THE DEVICE (table_plan_kernel, from bbox_kernel's box), so an update needs no host round
// trip; when the window or the table does not fit, every evaluation takes the exact path (same results, slower) and
// the host picks another kernel family for the next update from the plan summary it reads back asynchronously.
#pragma once
#include "mcl_kernels.cuh"

namespace mcl {

constexpr int kTabThreads = 1024;            // one CTA per SM: one K/T copy per SM
constexpr int kTabWarps = kTabThreads / 32;
constexpr int kTabQueue = 192;               // per-warp queue of deferred (lane, beam) entries
constexpr int kTabFixed = 129;               // T entries 0..128: the uniform classes
constexpr int kTabMaxBeams = 2047;           // queue entries are (lane << 11) | beam
constexpr int kTabApron = 4;                 // the class map covers the grid plus this many cells on every side
constexpr int kTabBatch = 4096;              // batch mode: consecutive particles that share one window (4 serpentine blocks
constexpr int kTabBatchSmall = 1024;         // of mcl_init_uniform), halved down to 1024 until the windows fit
constexpr float kTabB2 = 0.59033447f;        // 4/pi * atan(1/2): the octant boundaries in u8 units (see tab_sector)
constexpr float kTabS = 0.84697730f;         // 0.5 / kTabB2: u8 * S rounds to 0 inside +-B2, to +-1 beyond
constexpr float kTabC8 = 1.27323954f;        // 8 / (2 pi)

// per beam: interpolation ratio, beam angle, range in window-normalised units per axis (range * cells/m / (w - 1.5),
// ... / (h - 1.5)); the sector band d8 lives in a second array (one more broadcast LDS)
struct __align__(16) TabBeam { float ratio, theta, rcx, rcy; };

// Written by table_plan_kernel, read by score_table_kernel (device memory; a copy travels to pinned host memory as the
// host's hint for the next update).
struct TabPlan {
    int ok;                      // 0: the table pass is not applicable -> every evaluation takes the exact path
    int x0, y0, w, h;            // window, global cells
    int pitch_k;                 // K entries per row (2 bytes each, pitch_k / 2 odd; wide windows: 1 byte each, pitch_k / 4
                                 // odd): rows spread over the banks
    int cap_entries;             // T capacity, the fixed ones included
    unsigned off_k, off_t;       // byte offsets of K and T in dynamic shared memory
    int nseg;                    // wide windows: 64-cell segments per row (aligned in cell + bias_x)
    int seg_off;                 // wide windows: byte offset of the segment bases within a row of K
    // normalised coordinate of a window-relative coordinate c:  n = (c + off) * inv_s,  off = kappa - 0.5,
    // s = w - 1.5 (x) / h - 1.5 (y).  bits(1 + n) * mul = ((floor(c + kappa) + bias) << 32) | fraction(c + kappa) 2^32
    double off, inv_sx, inv_sy;
    unsigned mul_x, mul_y;       // (2w - 3) << 8, (2h - 3) << 8
    unsigned bias_x, bias_y;     // 127 w - 191, 127 h - 191
    unsigned frac_thr;           // ceil(2 kappa 2^32): fractions below it are within kappa of an integer
    float eps, kappa, t3;
    // particle validity (as FastPlan)
    float rho_lo, rho_hi, rho_abs, max_shift, coord_hi, reach, ang_room;
    float ulo_x, uhi_x, ulo_y, uhi_y;     // GLOBAL robot coordinates whose rays stay inside the cloud's bounding box +- reach
    unsigned hmin_x, hmin_y;              // EDGE 2: (cell + bias) values at or above these have a global cell >= 0
    float x2_lo_x, x2_lo_y;               // EDGE >= 1: normalised doubled endpoints at or above these are certainly >= 0
    int need_bytes;              // shared memory the window needs (K + fixed T + a minimum of entries)
    int reason;                  // why ok == 0 (diagnostic): 1 scan, 2 bbox, 3 window size, 4 eps, 5 disabled
    // batch mode (global localisation: the cloud as a whole does not fit one window): every batch of kTabBatch
    // consecutive particles gets a window of the SAME size (w, h) at its own origin (tab_batch_window), so everything
    // above that depends on the size only is shared and x0, y0, ulo/uhi, hmin, x2_lo are rewritten per batch
    int batch;                   // 1: this plan is a batch-mode plan
    int wide;                    // 1: one-byte class tile + segment bases (windows too large for 16-bit classes)
    int variant;                 // the variant this plan was made for (TabPlanIn::variant)
    int best;                    // the best applicable variant (-1: none): the host follows it on the next update
    int misfits;                 // batch mode: batches whose own window exceeds (w, h) -- they take the exact path
    // beam culling (one-window variants): where the robot can be (global cell coordinates) and which headings it can
    // have during the sweep, over every particle of the slice; table_cull_kernel turns that into a region per beam
    int cull_ok;
    double cull_x0, cull_x1, cull_y0, cull_y1, cull_a0, cull_a1;
    double rc6;                  // longest ray in cells + 6: the margin around a bounding box
    double x2_min;               // global coordinate the doubled endpoint must exceed (EDGE >= 1)
};

struct TabPlanIn {
    DevGrid grid;
    float max_range, min_range, max_abs_theta;
    double ratio_lo, ratio_hi;
    int num_beams;
    int scan_finite;
    int allow;
    int smem_total, smem_fixed;
    int variant;                 // kTabSingle16 .. kTabBatch8: what the host is going to launch
    int excluded;                // bit v: variant v is not to be used (its score table overflowed on this cloud)
    int cull;                    // beam culling allowed
    long long num_batches;
};

// Kernel variants, best first: one window for the whole slice or one per batch of particles; 16-bit classes (the class
// is the index of the score-table entry) or 8-bit classes with per-segment entry bases (half the shared memory per
// cell, six more instructions per evaluation).  A rebuild per batch costs less than the wide lookup.
constexpr int kTabSingle16 = 0, kTabSingle8 = 1, kTabBatch16 = 2, kTabBatch8 = 3;

__host__ __device__ inline long long tab_pitch_k(long long tw)
{
    long long p = (tw + 1) & ~1ll;
    if (((p >> 1) & 1) == 0) p += 2;
    return p;
}
__host__ __device__ inline long long tab_k_bytes(long long tw, long long th) { return (tab_pitch_k(tw) * 2 * th + 15) & ~15ll; }
// wide windows: a row holds one byte per cell, then one 16-bit entry base per 64-cell segment of the row (segments are
// aligned in (cell + bias_x), bias_x = 127 w - 191); rows of an odd number of words
__host__ __device__ inline long long tab_nseg(long long tw)
{
    const long long bias = 127 * tw - 191;
    return ((bias + tw - 1) >> 6) - (bias >> 6) + 1;
}
__host__ __device__ inline long long tab_seg_off(long long tw) { return (tw + 1) & ~1ll; }       // of the bases within a row
__host__ __device__ inline long long tab_pitch_k8(long long tw)
{
    long long p = (tab_seg_off(tw) + 2 * tab_nseg(tw) + 3) & ~3ll;
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}
__host__ __device__ inline long long tab_k8_bytes(long long tw, long long th) { return (tab_pitch_k8(tw) * th + 15) & ~15ll; }
// Batch mode sizes its windows so that T has room for one entry per 16 cells besides the fixed ones (free cells next to an
// occupied one: 2 % of the synthetic maps' cells, up to 6.5 % of a window's); a batch that needs more takes the exact path.
__host__ __device__ inline long long tab_batch_bytes(long long tw, long long th, bool wide)
{
    const long long e = tw * th / 16;
    return (wide ? tab_k8_bytes(tw, th) : tab_k_bytes(tw, th)) + 8 * (kTabFixed + (e > 256 ? e : 256));
}

// Unclipped cell range of the window of a bounding box (ordered-int floats, bbox of poses and parents):
// box +- (longest ray + 6 cells).  +6: every particle inside the box then passes make_tab_base's interior test, whose
// reach carries 4 cells of margin plus 1.5 for the window's border.  Returns 0, or the reason it has none (2: box not
// finite / empty, 3: out of the representable range).
__device__ __forceinline__ int tab_box_cells(int b0, int b1, int b2, int b3, const DevGrid& g, double rc6, long long& ux0,
                                             long long& uy0, long long& ux1, long long& uy1)
{
    auto unorder = [](int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); };
    const float mnx = unorder(b0), mny = unorder(b1), mxx = unorder(b2), mxy = unorder(b3);
    if (!(isfinite(mnx) && isfinite(mny) && isfinite(mxx) && isfinite(mxy)) || mnx > mxx || mny > mxy) return 2;
    const double cpm = (double)g.cells_per_meter;
    const double cx0 = floor(((double)mnx - (double)g.origin_x) * cpm - rc6);
    const double cy0 = floor(((double)mny - (double)g.origin_y) * cpm - rc6);
    const double cx1 = ceil(((double)mxx - (double)g.origin_x) * cpm + rc6);
    const double cy1 = ceil(((double)mxy - (double)g.origin_y) * cpm + rc6);
    if (!(fabs(cx0) < 1.0e6 && fabs(cy0) < 1.0e6 && cx1 - cx0 < 8192.0 && cy1 - cy0 < 8192.0)) return 3;
    ux0 = (long long)cx0; uy0 = (long long)cy0; ux1 = (long long)cx1; uy1 = (long long)cy1;
    return 0;
}

// The plan fields that depend on where the window lies: window [x0, x0 + w) x [y0, y0 + h) serving the particles whose
// rays stay inside the unclipped box [ux0, ux1] x [uy0, uy1].
__device__ __forceinline__ void tab_window_fields(TabPlan& pl, long long x0, long long y0, long long ux0, long long uy0,
                                                  long long ux1, long long uy1)
{
    pl.x0 = (int)x0; pl.y0 = (int)y0;
    pl.ulo_x = (float)((double)ux0 + 1.5 + (double)pl.reach); pl.uhi_x = (float)((double)(ux1 + 1) - 1.5 - (double)pl.reach);
    pl.ulo_y = (float)((double)uy0 + 1.5 + (double)pl.reach); pl.uhi_y = (float)((double)(uy1 + 1) - 1.5 - (double)pl.reach);
    pl.hmin_x = (unsigned)((long long)pl.bias_x - x0);
    pl.hmin_y = (unsigned)((long long)pl.bias_y - y0);
    pl.x2_lo_x = (float)((pl.x2_min - (double)x0 + pl.off) * pl.inv_sx * (1.0 + 1e-6) + 1e-7);
    pl.x2_lo_y = (float)((pl.x2_min - (double)y0 + pl.off) * pl.inv_sy * (1.0 + 1e-6) + 1e-7);
}

// Batch mode: places the plan's fixed-size window for the batch whose bounding box is bx.  The window is the batch's
// unclipped box shifted (not shrunk) into [-4, W + 3] x [-4, H + 3]: it still covers every cell of the box within four
// cells of the grid, and wherever rays can leave it its border lies four cells outside the grid (class 0).  false: the
// batch needs a larger window than the plan's (its evaluations take the exact path).
__device__ __forceinline__ bool tab_batch_window(TabPlan& pl, const int4 bx, const DevGrid& g)
{
    long long ux0, uy0, ux1, uy1;
    if (tab_box_cells(bx.x, bx.y, bx.z, bx.w, g, pl.rc6, ux0, uy0, ux1, uy1) != 0) return false;
    if (ux1 - ux0 + 1 > pl.w || uy1 - uy0 + 1 > pl.h) return false;
    const long long hx = (long long)g.width + 3 - (pl.w - 1), hy = (long long)g.height + 3 - (pl.h - 1);
    long long x0 = ux0 < hx ? ux0 : hx, y0 = uy0 < hy ? uy0 : hy;
    if (x0 < -4) x0 = -4;
    if (y0 < -4) y0 = -4;
    tab_window_fields(pl, x0, y0, ux0, uy0, ux1, uy1);
    return true;
}

// Bounding boxes: of the whole slice (box[0..3], ordered-int atomics) and of every batch of kTabBatch consecutive
// particles, kTabBatch or kTabBatchSmall (bboxes[b]).  For 16-bit classes, box[4], box[5] = the largest window width
// and height among the batches whose own window fits the shared-memory budget, box[6] = the number of batches that do
// not (non-finite poses, or spread too far); box[7], box[8] = the same maxima over the fitting batches that are at
// least as wide as tall, box[9] = the number of fitting batches that are taller than wide (the serpentine order of
// mcl_init_uniform runs along x: only the batches at its row turns are tall, and one window size for both orientations
// would have to be square).  box[10 .. 15] = the same for 8-bit classes.
// box[16], box[17] = the range of headings (ordered-int floats, relative to the parent heading of the slice's first
// particle, which goes to box[19]) the particles pass through between parent and pose; box[18] != 0: some heading is not
// finite or not wrapped.  (Beam culling, table_cull_kernel.)
__global__ void __launch_bounds__(256) table_bbox_kernel(const float* x, const float* y, const float* th, const float* px,
                                                         const float* py, const float* pth, long long lo, long long hi,
                                                         int batch, DevGrid grid, double rc6, long long k_budget, int* box,
                                                         int4* bboxes)
{
    __shared__ int red[6][8];
    const long long first = lo + (long long)blockIdx.x * batch;
    const long long last = first + batch < hi ? first + batch : hi;
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = (int)0x80000000, mxy = (int)0x80000000;
    int mna = 0x7fffffff, mxa = (int)0x80000000;
    bool bad_heading = false;
    const float th_ref = pth[lo];
    for (long long i = first + threadIdx.x; i < last; i += blockDim.x) {
        const int a = float_order(x[i]), b = float_order(y[i]), c = float_order(px[i]), d = float_order(py[i]);
        mnx = min(mnx, min(a, c)); mxx = max(mxx, max(a, c));
        mny = min(mny, min(b, d)); mxy = max(mxy, max(b, d));
        // the heading runs from the parent's to the pose's along the shorter way round (interpolation.hpp:41)
        const float ta = th[i], tb = pth[i];
        if (!(fabsf(ta) <= 3.15f && fabsf(tb) <= 3.15f && fabsf(th_ref) <= 3.15f)) bad_heading = true;
        const double db = fold_pi(__dsub_rn((double)tb, (double)th_ref));
        const double da = db + fold_pi(__dsub_rn((double)ta, (double)tb));
        const float lo_f = __double2float_rd(fmin(db, da)), hi_f = __double2float_ru(fmax(db, da));
        mna = min(mna, float_order(lo_f)); mxa = max(mxa, float_order(hi_f));
    }
    for (int off = 16; off > 0; off >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, off));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, off));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, off));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, off));
        mna = min(mna, __shfl_xor_sync(0xffffffffu, mna, off));
        mxa = max(mxa, __shfl_xor_sync(0xffffffffu, mxa, off));
    }
    if (__any_sync(0xffffffffu, bad_heading) && (threadIdx.x & 31) == 0) atomicOr(box + 18, 1);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = mnx; red[1][threadIdx.x >> 5] = mny;
        red[2][threadIdx.x >> 5] = mxx; red[3][threadIdx.x >> 5] = mxy;
        red[4][threadIdx.x >> 5] = mna; red[5][threadIdx.x >> 5] = mxa;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            mnx = min(mnx, red[0][k]); mny = min(mny, red[1][k]); mxx = max(mxx, red[2][k]); mxy = max(mxy, red[3][k]);
            mna = min(mna, red[4][k]); mxa = max(mxa, red[5][k]);
        }
        atomicMin(box + 0, mnx); atomicMin(box + 1, mny);
        atomicMax(box + 2, mxx); atomicMax(box + 3, mxy);
        atomicMin(box + 16, mna); atomicMax(box + 17, mxa);
        if (blockIdx.x == 0) box[19] = __float_as_int(th_ref);
        if (bboxes) {
            bboxes[blockIdx.x] = make_int4(mnx, mny, mxx, mxy);
            long long ux0, uy0, ux1, uy1;
            const bool cells = tab_box_cells(mnx, mny, mxx, mxy, grid, rc6, ux0, uy0, ux1, uy1) == 0;
            const int bw = (int)(ux1 - ux0 + 1), bh = (int)(uy1 - uy0 + 1);
            for (int wide = 0; wide < 2; ++wide) {
                int* bx = box + 4 + 6 * wide;
                if (cells && tab_batch_bytes(bw, bh, wide != 0) <= k_budget) {
                    atomicMax(bx + 0, bw); atomicMax(bx + 1, bh);
                    if (bw >= bh) { atomicMax(bx + 3, bw); atomicMax(bx + 4, bh); }
                    else atomicAdd(bx + 5, 1);
                } else {
                    atomicAdd(bx + 2, 1);
                }
            }
        }
    }
}

// One thread: boxes -> window(s) -> budget.  Re-arms the boxes for the next table_bbox_kernel.
// Also re-arms what the score kernel accumulates into: the build summary and the deferred / gather counters.
__global__ void table_plan_kernel(const TabPlanIn in, int* box, TabPlan* out, int* build, unsigned long long* deferred,
                                  unsigned long long* gathers)
{
    for (int i = 0; i < 8; ++i) build[i] = 0;
    *deferred = 0ull;
    if (gathers) *gathers = 0ull;
    TabPlan pl;
    memset(&pl, 0, sizeof(pl));
    const int b0 = box[0], b1 = box[1], b2 = box[2], b3 = box[3];
    long long bdim[2][3];            // [wide]: window width, height, upper bound of the batches that will not fit it
    const int ha0 = box[16], ha1 = box[17], hbad = box[18];
    const float th_ref = __int_as_float(box[19]);
    box[16] = 0x7fffffff; box[17] = (int)0x80000000; box[18] = 0;
    box[0] = 0x7fffffff; box[1] = 0x7fffffff; box[2] = (int)0x80000000; box[3] = (int)0x80000000;
    {
        const long long room_b = (long long)in.smem_total - in.smem_fixed - 64;
        for (int wide = 0; wide < 2; ++wide) {
            const int* bx = box + 4 + 6 * wide;
            bdim[wide][0] = bx[0]; bdim[wide][1] = bx[1]; bdim[wide][2] = bx[2];
            // both orientations in one window size if that fits, else the wide batches' size (the tall ones: exact path)
            if (bx[0] >= 8 && bx[1] >= 8 && tab_batch_bytes(bx[0], bx[1], wide != 0) > room_b && bx[3] >= 8 && bx[4] >= 8) {
                bdim[wide][0] = bx[3]; bdim[wide][1] = bx[4]; bdim[wide][2] = (long long)bx[2] + bx[5];
            }
        }
    }
    for (int i = 4; i < 16; ++i) box[i] = 0;
    const double cpm = (double)in.grid.cells_per_meter;
    const double Rc = (double)in.max_range * cpm;
    const double rho_max = fmax(fabs(in.ratio_lo), fabs(in.ratio_hi));
    pl.rc6 = Rc + 6.0;
    pl.variant = in.variant;
    pl.best = -1;
    do {
        if (!in.allow) { pl.reason = 5; break; }
        if (!in.scan_finite || in.num_beams < 1 || in.num_beams > kTabMaxBeams || !isfinite(Rc) || !(cpm > 0.0) ||
            !((double)in.min_range * cpm >= 2.5) || !(in.max_abs_theta <= 6.3f) || !(in.ratio_lo >= -1.0) ||
            !(in.ratio_hi <= 2.0)) { pl.reason = 1; break; }
        // one window: the slice's box, not clipped to the grid for the particle tests (cells outside read 0) ...
        long long ux0 = 0, uy0 = 0, ux1 = 0, uy1 = 0;
        const int br = tab_box_cells(b0, b1, b2, b3, in.grid, pl.rc6, ux0, uy0, ux1, uy1);
        if (br == 2) { pl.reason = 2; break; }
        // ... and clipped to the grid plus four cells for the tile: everything beyond is class 0, and so are the clipped
        // window's border cells, onto which the saturating coordinate arithmetic maps rays that leave it
        const long long x0 = ux0 > -4 ? ux0 : -4, y0 = uy0 > -4 ? uy0 : -4;
        const long long x1 = ux1 < in.grid.width + 3 ? ux1 : in.grid.width + 3, y1 = uy1 < in.grid.height + 3 ? uy1 : in.grid.height + 3;
        const long long sw = x1 - x0 + 1, sh = y1 - y0 + 1;
        const long long room_base = (long long)in.smem_total - in.smem_fixed - 64;
        const long long min_t = 8 * (kTabFixed + 256);
        bool fit[4];
        const bool single = br == 0 && sw >= 8 && sh >= 8;              // (sw < 8: a cloud whose rays cannot reach the grid)
        fit[kTabSingle16] = single && room_base - tab_k_bytes(sw, sh) >= min_t;
        fit[kTabSingle8] = single && room_base - tab_k8_bytes(sw, sh) >= min_t;
        for (int wide = 0; wide < 2; ++wide)
            fit[kTabBatch16 + wide] = in.num_batches > 0 && bdim[wide][0] >= 8 && bdim[wide][1] >= 8 &&
                                      bdim[wide][2] * 50 <= in.num_batches &&
                                      room_base - tab_batch_bytes(bdim[wide][0], bdim[wide][1], wide != 0) >= 0;
        for (int v = 3; v >= 0; --v)
            if (fit[v] && !((in.excluded >> v) & 1)) pl.best = v;
        pl.need_bytes = br == 0 ? (int)fmin(2.0e9, (double)(in.smem_fixed + tab_k_bytes(sw, sh) + 64 + min_t)) : 0;
        const int v = in.variant;
        if (v < 0 || v > 3 || !fit[v] || ((in.excluded >> v) & 1)) { pl.reason = 3; break; }
        const bool batch = v >= kTabBatch16, wide = (v & 1) != 0;
        const long long tw = batch ? bdim[wide][0] : sw, th = batch ? bdim[wide][1] : sh;
        pl.misfits = batch ? (int)bdim[wide][2] : 0;
        const long long pitch_k = wide ? tab_pitch_k8(tw) : tab_pitch_k(tw);
        const long long k_bytes = wide ? tab_k8_bytes(tw, th) : tab_k_bytes(tw, th);
        const long long room = room_base - k_bytes;
        // error budget (cells).  Reference vs the real-valued model: as fast_plan (mcl_engine.cu).  Float model vs the
        // same: roundings of dS, rho and rc, the angle roundings and the measured SFU error; its coordinate roundings
        // (the normalised robot coordinate, the ratio FFMA, the endpoint FFMA, the +1.0) are below 2^-24 of the
        // window's extent each.
        const double u = 5.9604644775390625e-08;
        const double Cm = (double)(ux1 > uy1 ? ux1 : uy1) + 2.0;
        if (br != 0 || Cm > 16000.0) { pl.reason = 3; pl.best = -1; break; }
        const double Xm = Cm / cpm + fmax(fabs((double)in.grid.origin_x), fabs((double)in.grid.origin_y));
        const double max_shift = 64.0;
        const double Ce = Cm + Rc;
        const double e_ref = cpm * u * Xm + 2.0 * u * Ce + 2.0 * u * Rc + Rc * (20.0 * u + 1.2e-7) + 1e-9;
        // (angle: roundings of theta_r, of the folded beam angle and of their difference, + the SFU's error)
        const double e_apx = 6.0 * u * (double)(tw > th ? tw : th) + (1.0 + 2.0 * rho_max) * u * max_shift + 3.0 * u * Rc +
                             Rc * ((3.14159265358979 * (3.0 * rho_max + 2.0) + 9.5) * u + (double)kFastTrigErr);
        const double eps = 1.25 * (e_ref + e_apx) + 1e-6;
        const double kappa = eps + 1e-5;
        if (kappa > 1.0 / 16.0) { pl.reason = 4; pl.best = -1; break; }
        pl.ok = 1;
        pl.batch = batch ? 1 : 0;
        pl.wide = wide ? 1 : 0;
        pl.w = (int)tw; pl.h = (int)th; pl.pitch_k = (int)pitch_k;
        pl.nseg = (int)tab_nseg(tw);
        pl.seg_off = (int)tab_seg_off(tw);
        pl.cap_entries = (int)(room / 8);
        pl.off_k = (unsigned)in.smem_fixed;
        pl.off_t = (unsigned)(in.smem_fixed + k_bytes);
        pl.off = kappa - 0.5;
        pl.inv_sx = 1.0 / ((double)tw - 1.5); pl.inv_sy = 1.0 / ((double)th - 1.5);
        pl.mul_x = (unsigned)(2 * tw - 3) << 8; pl.mul_y = (unsigned)(2 * th - 3) << 8;
        pl.bias_x = (unsigned)(127 * tw - 191); pl.bias_y = (unsigned)(127 * th - 191);
        pl.frac_thr = (unsigned)ceil(2.0 * kappa * 4294967296.0);
        pl.eps = (float)eps; pl.kappa = (float)kappa;
        pl.t3 = (float)(3.0 * (1.0 + 2.0 * eps) + 4.0 * u * Rc + 1e-4);
        pl.rho_lo = (float)in.ratio_lo; pl.rho_hi = (float)in.ratio_hi; pl.rho_abs = (float)(rho_max * (1.0 + 1e-6));
        pl.max_shift = (float)max_shift;
        pl.coord_hi = (float)(Cm - 1.0);
        pl.reach = (float)(Rc * (1.0 + 1e-6) + 4.0);
        pl.ang_room = 9.5f - fminf(in.max_abs_theta, 3.1415928f);      // kFastTrigErr is measured for |angle| <= 9.5
        pl.x2_min = 3.0 * eps + 2.0 * kappa + 1e-3;        // global coordinate the doubled endpoint must exceed
        tab_window_fields(pl, x0, y0, ux0, uy0, ux1, uy1); // (batch mode: rewritten for every batch)
        // beam culling: one window, interpolation ratios within [0, 1] (the ray origin then lies between parent and
        // pose), wrapped finite headings spanning less than 2.5 rad
        auto unorder = [](int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); };
        const float a0 = unorder(ha0), a1 = unorder(ha1);
        if (!batch && in.cull && !hbad && in.ratio_lo >= 0.0 && in.ratio_hi <= 1.0 && isfinite(a0) && isfinite(a1) && a0 <= a1 &&
            a1 - a0 < 2.5f) {
            pl.cull_ok = 1;
            pl.cull_x0 = ((double)unorder(b0) - (double)in.grid.origin_x) * cpm; pl.cull_x1 = ((double)unorder(b2) - (double)in.grid.origin_x) * cpm;
            pl.cull_y0 = ((double)unorder(b1) - (double)in.grid.origin_y) * cpm; pl.cull_y1 = ((double)unorder(b3) - (double)in.grid.origin_y) * cpm;
            pl.cull_a0 = (double)th_ref + (double)a0; pl.cull_a1 = (double)th_ref + (double)a1;
        }
    } while (false);
    *out = pl;
}

// BEAM CULLING.  A beam whose endpoint cell is class 0 for EVERY particle of the slice scores 0 for every particle (no
// positive cell within two cells of the endpoint cell, hence none among the reference's three reads), so the table pass
// need not evaluate it at all -- typically the rays that ended in open space at the sensor's maximum range.  One CTA
// per beam bounds where its endpoints can lie.  The ray origin S lies in the bounding box of poses and parents (centre
// C, half extents hx, hy), the heading in [a0, a1], so with am the middle direction of the beam, hw the half width of
// its directions, and (u, v) the coordinates along / across am relative to C:
//     u in [rc cos hw - bu, rc + bu],   |v| <= rc sin hw + bv,     bu = hx |cos am| + hy |sin am|, bv = hx |sin am| + hy |cos am|
// (an annular sector swept by the box, bounded by an oriented rectangle).  + 4.5 cells: 3 for the roundings of either
// model and the reference's truncation, 1.5 because a cell is tested at its corner.  The beam is culled when every cell
// whose corner lies in that rectangle is class 0 in the class map (cells beyond the map's apron are: they lie four or
// more cells outside the grid; cells with a negative coordinate inside the apron are class 1, so truncation toward zero
// never matters).  flags[b] = 1: culled.  count[0] += culled beams, count[1] = 1 (the pass ran).
__global__ void __launch_bounds__(256) table_cull_kernel(const TabPlan* plan, const Beam* beams, int num_beams, DevGrid grid,
                                                         const uint8_t* cls, int cpitch, uint8_t* flags, int* count)
{
    const int b = blockIdx.x;
    if (b >= num_beams) return;
    const TabPlan& pl = *plan;
    if (b == 0 && threadIdx.x == 0) count[1] = 1;
    if (!pl.ok || !pl.cull_ok) {
        if (threadIdx.x == 0) flags[b] = 0;
        return;
    }
    const Beam bm = beams[b];
    const double rc = (double)__fmul_rn(bm.range, grid.cells_per_meter);
    const double am = 0.5 * (pl.cull_a0 + pl.cull_a1) - (double)bm.theta, hw = 0.5 * (pl.cull_a1 - pl.cull_a0) * (1.0 + 1e-6) + 1e-6;
    double sn, cs;
    sincos(am, &sn, &cs);
    const double cx = 0.5 * (pl.cull_x0 + pl.cull_x1), cy = 0.5 * (pl.cull_y0 + pl.cull_y1);
    const double hx = 0.5 * (pl.cull_x1 - pl.cull_x0), hy = 0.5 * (pl.cull_y1 - pl.cull_y0);
    const double bu = hx * fabs(cs) + hy * fabs(sn), bv = hx * fabs(sn) + hy * fabs(cs);
    const double m = 4.5;
    const double u0 = rc * cos(hw) * (1.0 - 1e-6) - bu - m, u1 = rc * (1.0 + 1e-6) + bu + m, vm = rc * sin(hw) * (1.0 + 1e-6) + bv + m;
    // axis-aligned bounds of the oriented rectangle
    const double ex = fmax(fabs(u0), fabs(u1)) * fabs(cs) + vm * fabs(sn), ey = fmax(fabs(u0), fabs(u1)) * fabs(sn) + vm * fabs(cs);
    const double fx0 = floor(cx - ex), fx1 = ceil(cx + ex), fy0 = floor(cy - ey), fy1 = ceil(cy + ey);
    bool nonzero = !(fabs(fx0) < 1.0e6 && fabs(fy0) < 1.0e6 && fabs(fx1) < 1.0e6 && fabs(fy1) < 1.0e6);
    if (!nonzero) {
        const int x0 = max((int)fx0, -kTabApron), x1 = min((int)fx1, grid.width + kTabApron - 1);
        const int y0 = max((int)fy0, -kTabApron), y1 = min((int)fy1, grid.height + kTabApron - 1);
        const int w = x1 - x0 + 1, h = y1 - y0 + 1;
        if (w > 0 && h > 0) {
            if ((long long)w * h > 262144) nonzero = true;         // (not worth scanning: a wide heading range)
            else {
                const float fcs = (float)cs, fsn = (float)sn, fu0 = (float)u0 - 0.01f, fu1 = (float)u1 + 0.01f, fvm = (float)vm + 0.01f;
                const float ox = (float)((double)x0 - cx), oy = (float)((double)y0 - cy);
                unsigned acc = 0u;
                for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
                    const int ry = i / w, rx = i - ry * w;
                    const float dx = ox + (float)rx, dy = oy + (float)ry;
                    const float u = dx * fcs + dy * fsn, v = dy * fcs - dx * fsn;
                    if (u >= fu0 && u <= fu1 && fabsf(v) <= fvm)
                        acc |= (unsigned)__ldg(cls + (size_t)(y0 + ry + kTabApron) * cpitch + (x0 + rx + kTabApron));
                }
                nonzero = acc != 0u;
            }
        }
    }
    const int any = __syncthreads_or(nonzero ? 1 : 0);
    if (threadIdx.x == 0) {
        flags[b] = any ? 0 : 1;
        if (!any) atomicAdd(count, 1);
    }
}

// Per-particle constants of the table pass: robot coordinate at rho = 0 in window-normalised units, its change over
// rho = 0..1, heading likewise.
struct TabBase {
    float sxn, syn, thb, dsxn, dsyn, dth;
    bool ok;
    int edge;       // 0: neither endpoints nor doubled endpoints can have a negative global coordinate;
                    // 1: doubled endpoints may; 2: endpoints may too
};

template <bool INTERP>
__device__ __forceinline__ TabBase make_tab_base(float xa, float ya, float tha, float xb, float yb, float thb, double gx,
                                                 double gy, double cpm_d, const TabPlan& pl)
{
    TabBase f;
    double gsx, gsy, ddx = 0.0, ddy = 0.0;      // robot cell coordinate (global, double) and its change over the sweep
    if (INTERP) {
        gsx = __dmul_rn(__dsub_rn((double)xb, gx), cpm_d);
        gsy = __dmul_rn(__dsub_rn((double)yb, gy), cpm_d);
        f.thb = thb;
        ddx = __dmul_rn((double)__fsub_rn(xa, xb), cpm_d);              // the reference's float difference (interpolation.hpp:39)
        ddy = __dmul_rn((double)__fsub_rn(ya, yb), cpm_d);
        f.dth = (float)fold_pi(__dsub_rn((double)tha, (double)thb));    // angle_diff (:41)
    } else {
        gsx = __dmul_rn(__dsub_rn((double)xa, gx), cpm_d);
        gsy = __dmul_rn(__dsub_rn((double)ya, gy), cpm_d);
        f.thb = tha;
        f.dth = 0.0f;
    }
    f.sxn = (float)__dmul_rn(__dadd_rn(__dsub_rn(gsx, (double)pl.x0), pl.off), pl.inv_sx);
    f.syn = (float)__dmul_rn(__dadd_rn(__dsub_rn(gsy, (double)pl.y0), pl.off), pl.inv_sy);
    f.dsxn = (float)__dmul_rn(ddx, pl.inv_sx);
    f.dsyn = (float)__dmul_rn(ddy, pl.inv_sy);
    const float gxb = (float)gsx, gyb = (float)gsy, dsx = (float)ddx, dsy = (float)ddy;
    const float xs0 = __fmaf_rn(dsx, pl.rho_lo, gxb), xs1 = __fmaf_rn(dsx, pl.rho_hi, gxb);
    const float ys0 = __fmaf_rn(dsy, pl.rho_lo, gyb), ys1 = __fmaf_rn(dsy, pl.rho_hi, gyb);
    const float xlo = fminf(xs0, xs1), xhi = fmaxf(xs0, xs1), ylo = fminf(ys0, ys1), yhi = fmaxf(ys0, ys1);
    const float lo = fminf(xlo, ylo), hi = fmaxf(xhi, yhi);
    // every ray of the particle must end inside the cloud's bounding box +- reach (always true for interpolation ratios
    // in [0, 1]; extrapolating ratios can leave it): beyond it the window's border need not be class 0
    // (NaN anywhere: the comparisons fail and the particle is left to the exact path)
    const bool in_uwin = xlo >= pl.ulo_x && xhi <= pl.uhi_x && ylo >= pl.ulo_y && yhi <= pl.uhi_y;
    f.ok = lo >= 1.0f && hi <= pl.coord_hi && fabsf(dsx) <= pl.max_shift && fabsf(dsy) <= pl.max_shift &&
           fabsf(f.thb) <= 3.15f && fabsf(f.dth) <= 3.15f &&
           __fmaf_rn(pl.rho_abs, fabsf(f.dth), fabsf(f.thb)) <= pl.ang_room && in_uwin;
    f.edge = lo >= 2.0f * pl.reach ? 0 : (lo >= pl.reach ? 1 : 2);
    return f;
}

// Hot-loop constants of one CTA (registers).
struct TabConst {
    unsigned mul_x, mul_y, frac_thr;
    unsigned pitch2;        // bytes per K row
    unsigned kbase;         // shared address of K minus the bias of both axes
    unsigned k0;            // shared address of T / 8: the K tile stores class + k0
    unsigned k129;          // k0 + kTabFixed
    unsigned lf;            // likelihood-field mode (gather counting only)
    unsigned sdelta;        // WIDE: from the row term of a cell's address to the row's segment bases (minus the bias)
};

__device__ __forceinline__ float fma_sat(float a, float b, float c)
{
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));     // clamps to [0, 1]; NaN -> +0
    return r;
}

// One certified evaluation of the table pass.  Returns true when certain; add = its score in half units (else 0).
// EDGE (warp-uniform, from TabBase::edge): 1 adds the test that the doubled endpoint has no negative coordinate (the
// sector is not certified otherwise), 2 also the test that the endpoint's global cell is >= 0.
// WIDE: the class tile holds one byte per cell -- 0, 1, 2..128 as above, 129 + i = the i-th cell with a table entry of
// its 64-cell row segment -- and a second lookup (one 16-bit base per segment) completes the entry's index.
template <bool INTERP, bool COUNT, int EDGE, bool WIDE>
__device__ __forceinline__ bool tab_eval(const TabBase& p, const TabBeam& b, float d8, const TabConst& tc,
                                         const TabPlan& pl, int& add, int& gathers)
{
    const float sxn = INTERP ? __fmaf_rn(p.dsxn, b.ratio, p.sxn) : p.sxn;
    const float syn = INTERP ? __fmaf_rn(p.dsyn, b.ratio, p.syn) : p.syn;
    const float thr = INTERP ? __fmaf_rn(p.dth, b.ratio, p.thb) : p.thb;
    const float a = __fsub_rn(thr, b.theta);
    const float s = __sinf(a), c = __cosf(a);
    // normalised endpoint, clamped onto the window; 1 + n has a linear mantissa
    const float nx = fma_sat(b.rcx, c, sxn), ny = fma_sat(b.rcy, s, syn);
    const float bxf = __fadd_rn(nx, 1.0f), byf = __fadd_rn(ny, 1.0f);
    // bits * (2w - 3) 2^8: high word = cell + bias, low word = fraction * 2^32 (both of coordinate + kappa)
    unsigned fx, hx, fy, hy, kaddr;
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(fx), "=r"(hx) : "r"(__float_as_uint(bxf)), "r"(tc.mul_x));
    asm("{\n\t.reg .b64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(fy), "=r"(hy) : "r"(__float_as_uint(byf)), "r"(tc.mul_y));
    bool cell_ok = min(fx, fy) >= tc.frac_thr;
    if (EDGE >= 2) cell_ok = cell_ok & (hx >= pl.hmin_x) & (hy >= pl.hmin_y);       // global cell >= 0 on both axes
    unsigned krow;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(krow) : "r"(hy), "r"(tc.pitch2), "r"(tc.kbase));
    if (WIDE) kaddr = krow + hx;
    else asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(kaddr) : "r"(hx), "r"(krow));
    bool x2_pos = true;
    if (EDGE >= 1)      // the doubled endpoint must not have a negative coordinate (the reference truncates toward zero there)
        x2_pos = (__fmaf_rn(b.rcx, c, nx) >= pl.x2_lo_x) & (__fmaf_rn(b.rcy, s, ny) >= pl.x2_lo_y);
    unsigned K, sb = 0u;
    if (WIDE) {
        unsigned saddr;
        asm("ld.shared.u8 %0, [%1];" : "=r"(K) : "r"(kaddr));
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(saddr) : "r"(hx >> 6), "r"(krow + tc.sdelta));
        asm("ld.shared.u16 %0, [%1];" : "=r"(sb) : "r"(saddr));
    } else {
        asm("ld.shared.u16 %0, [%1];" : "=r"(K) : "r"(kaddr));
    }
    // sector: u8 = 8t - 2 round(4t) in [-1, 1] (t = a / 2pi) is the offset from the nearest axis, +-1 = 45 degrees;
    // n = 2 round(4t) + round(u8 * S) with S = 0.5 / B2, so the rounding flips exactly at the octant boundaries +-B2
    const float r8 = __fmaf_rn(a, kTabC8, 25165824.0f);                 // 1.5 * 2^24: ulp 2
    const float u8 = __fmaf_rn(a, kTabC8, -__fsub_rn(r8, 25165824.0f));
    const float wq = __fmaf_rn(u8, kTabS, __fsub_rn(r8, 12582912.0f));   // 1.5 * 2^23 + 2 round(4t): ulp 1
    // K carries the table's base (shared address / 8), so K * 8 + sector is the address of the score byte
    unsigned v, taddr;
    const bool dir_ok = x2_pos & (fabsf(__fsub_rn(fabsf(u8), kTabB2)) > d8);
    bool certain;
    if (WIDE) {
        // entry = k0 + K for the uniform classes, base of the segment (which carries k0 - 129) + K for the others
        const bool uni = K < (unsigned)kTabFixed;
        const unsigned e = K + (uni ? tc.k0 : sb);
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(taddr) : "r"(e), "r"(__float_as_uint(wq) & 7u));
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(taddr));
        certain = (K == 0u) | (cell_ok & (uni | dir_ok));
        if (COUNT) gathers += certain ? ((K - 2u < 127u) | (tc.lf != 0u) ? 1 : 3) : 0;
    } else {
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(taddr) : "r"(K), "r"(__float_as_uint(wq) & 7u));
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(taddr));
        certain = (K == tc.k0) | (cell_ok & ((K < tc.k129) | dir_ok));
        if (COUNT) gathers += certain ? ((K - tc.k0 - 2u < 127u) | (tc.lf != 0u) ? 1 : 3) : 0;
    }
    add = certain ? (int)v : 0;
    return certain;
}

// How the kernel deals its work (shared with the host, which zeroes the scores of the split units): nunits units of 32
// particles over `slots` warps; the units of the last, partial round are split into `split` beam ranges each.
__host__ __device__ inline void table_tail_split(long long nunits, long long slots, int nwords, long long& whole_units,
                                                 int& split, long long& items)
{
    const long long rem = nunits % slots;
    whole_units = nunits - rem;
    split = 1;
    if (rem > 0) {
        long long s = slots / rem;
        if (s > nwords) s = nwords;
        if (s > 1) split = (int)s;
    }
    if (split == 1) whole_units = nunits;
    items = whole_units + (nunits - whole_units) * split;
}

struct TabArgs {
    const float *x, *y, *th, *px, *py, *pth;
    int32_t* score2;
    long long lo, hi;
    const Beam* beams;
    int num_beams;
    DevGrid grid;                   // the mirror itself (exact path, table build)
    const int8_t* fast_cells;       // derived map: -1 = nothing positive within two cells
    const int8_t* lf_cells;         // likelihood-field mode (sensor_mode 1): the field u (0..127), same pitch; else null
    const TabPlan* plan;
    unsigned long long* gather_counter;
    unsigned long long* deferred_counter;
    int* build_info;                // [0] table entries the window needs (max over CTAs), [1] CTAs (batch mode: batches)
                                    // that overflowed, [2] batch mode: the next batch to take, [3] batches without a
                                    // window, [4] particles outside the table pass's domain (diagnostics), [5] beams
                                    // culled, [6] the culling pass ran
    const uint8_t* cls;             // class map (derive_class_map_kernel) ...
    const unsigned long long* pack; // ... and the score-table entries of its class-255 cells
    int cpitch;                     // bytes per row of the class map
    const int4* bboxes;             // batch mode: bounding box of every batch (table_bbox_kernel)
    int batch;                      // batch mode: particles per batch
    const uint8_t* cull;            // one-window variants: beams the table pass skips (table_cull_kernel), else null
};

// Cold-path state shared by the CTA (static shared memory): what the exact evaluations need.
struct TabCold {
    DevGrid grid;
    const Beam* sbeams;
    const uint16_t* ktile;      // class tile (null: not built -- read the mirror); wide windows: bytes
    int x0, y0, w, h, pitch_k;
    unsigned k0;
    int wide;
    int lf;                     // likelihood-field mode: a ray scores 2 u(endpoint cell), no neighbour steps
    unsigned long long deferred, gathers;
};

// Cell value as the score sees it: the log-odds where positive, else 0 (sensor_model.cpp:41-57 only ever tests
// "> 0").  Inside the window the class tile has it (classes 2..128 = occupied, log-odds class - 1); outside, the mirror.
__device__ __forceinline__ int tab_cell_value(const TabCold* c, int gx, int gy)
{
    const int tx = (int)((unsigned)gx - (unsigned)c->x0), ty = (int)((unsigned)gy - (unsigned)c->y0);
    if (c->ktile && (unsigned)tx < (unsigned)c->w && (unsigned)ty < (unsigned)c->h) {
        const unsigned cls = c->wide ? (unsigned)reinterpret_cast<const uint8_t*>(c->ktile)[ty * c->pitch_k + tx]
                                     : (unsigned)c->ktile[ty * c->pitch_k + tx] - c->k0;
        return (cls - 2u < 127u) ? (int)cls - 1 : 0;
    }
    return max(grid_read(c->grid, gx, gy), 0);
}

// Exact evaluation of one queue entry per lane: the literal restatement (exact_endpoint + the integer rules of
// sensor_model.cpp:41-86 with x86 conversion semantics), cell values from the class tile.
template <bool INTERP, bool COUNT>
__device__ __noinline__ void tab_drain_round(unsigned entry, bool active, float xa, float ya, float tha, float xb,
                                             float yb, float thb, TabCold* cold, int* wacc)
{
    const int src = (int)(entry >> 11) & 31, j = (int)(entry & 2047u);
    const float pxa = __shfl_sync(0xffffffffu, xa, src), pya = __shfl_sync(0xffffffffu, ya, src);
    const float ptha = __shfl_sync(0xffffffffu, tha, src), pxb = __shfl_sync(0xffffffffu, xb, src);
    const float pyb = __shfl_sync(0xffffffffu, yb, src), pthb = __shfl_sync(0xffffffffu, thb, src);
    if (active) {
        GridConst gc;
        gc.gx = (double)cold->grid.origin_x; gc.gy = (double)cold->grid.origin_y;
        gc.cpm = cold->grid.cells_per_meter; gc.cpm_d = (double)cold->grid.cells_per_meter;
        gc.trig.hpi_inv = gs_k[0]; gc.trig.hpi = gs_k[1]; gc.trig.s1 = gs_k[2]; gc.trig.s2 = gs_k[3]; gc.trig.s3 = gs_k[4];
        gc.trig.c0 = gs_k[5]; gc.trig.c1 = gs_k[6]; gc.trig.c2 = gs_k[7]; gc.trig.c3 = gs_k[8]; gc.trig.c4 = gs_k[9];
        const RayBase rb = make_ray_base(pxa, pya, ptha, pxb, pyb, pthb);
        float sx, sy, px, py, e1x, e1y;
        exact_endpoint<INTERP>(rb, cold->sbeams[j], gc, sx, sy, px, py, e1x, e1y);
        const int ex = f2i_x86(e1x), ey = f2i_x86(e1y);                                   // sensor_model.cpp:34-35
        int v, g = 1;
        const int odds = tab_cell_value(cold, ex, ey);                                    // :41
        if (odds > 0 || cold->lf) {
            v = 2 * odds;
        } else {
            const int xx = f2i_x86(__fadd_rn(__fmul_rn(2.0f, px), sx));                   // :37-38
            const int xy = f2i_x86(__fadd_rn(__fmul_rn(2.0f, py), sy));
            int ax, ay, bx, by;
            bresenham_step(ex, ey, f2i_x86(sx), f2i_x86(sy), ax, ay);                     // :48
            bresenham_step(ex, ey, xx, xy, bx, by);                                       // :49
            const int o1 = tab_cell_value(cold, ax, ay), o2 = tab_cell_value(cold, bx, by);
            v = o1 > 0 ? o1 : o2;
            g = 3;
        }
        if (v) atomicAdd(wacc + src, v);
        if (COUNT) atomicAdd(&cold->gathers, (unsigned long long)g);
    }
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, active));
    if ((threadIdx.x & 31) == 0) atomicAdd(&cold->deferred, (unsigned long long)cnt);
}

// Hands the set bits of the warp's mask words (m0..m3: beams 32 wbase .. 32 wbase + 127 of each lane's particle) to the
// exact path: the lanes append their (lane, beam) entries to the warp's queue in lane order (positions from a warp scan
// of the pop-counts), the queue is drained 32 entries at a time, and whatever did not fit goes in the next pass -- so
// every drain round has 32 active lanes whatever the number and distribution of uncertain beams (a particle outside
// the table pass's domain has all of them set).  Returns the new queue length (< 32).
template <bool INTERP, bool COUNT>
__device__ __noinline__ int tab_defer_words(unsigned m0, unsigned m1, unsigned m2, unsigned m3, int wbase, int qn,
                                            float xa, float ya, float tha, float xb, float yb, float thb, TabCold* cold,
                                            uint16_t* q, int* wacc)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        const int c = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) break;
        int pos = qn + incl - c;
        int beam0 = wbase * 32;
#pragma unroll
        for (int part = 0; part < 4; ++part) {
            unsigned m = part == 0 ? m0 : (part == 1 ? m1 : (part == 2 ? m2 : m3));
            while (m && pos < kTabQueue) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                q[pos++] = (uint16_t)((lane << 11) | (beam0 + k));
            }
            if (part == 0) m0 = m; else if (part == 1) m1 = m; else if (part == 2) m2 = m; else m3 = m;
            beam0 += 32;
        }
        qn = min(qn + total, kTabQueue);
        __syncwarp();
        while (qn >= 32) {
            qn -= 32;
            tab_drain_round<INTERP, COUNT>(q[qn + lane], true, xa, ya, tha, xb, yb, thb, cold, wacc);
        }
        __syncwarp();
    }
    return qn;
}

#ifndef MCL_TAB_UNROLL
#define MCL_TAB_UNROLL 4
#endif

// Class of one window cell (see the head of this file): 0, 1, 2..128, or kTabNeedsEntry with the cell's eight
// neighbours in n[] (sector order) when the cell gets a score-table entry.
constexpr unsigned kTabNeedsEntry = 0xffffffffu;
__device__ __forceinline__ unsigned tab_classify(const TabArgs& a, int gxc, int gyc, int n[8])
{
    const int W = a.grid.width, H = a.grid.height, gp = a.grid.pitch;
    const int8_t* raw = a.grid.cells;
    auto rawc = [&](int cx, int cy) -> int {
        return ((unsigned)cx < (unsigned)W && (unsigned)cy < (unsigned)H) ? (int)__ldg(raw + (size_t)cy * gp + cx) : 0;
    };
    if (gxc <= -4 || gyc <= -4 || gxc >= W + 3 || gyc >= H + 3)
        return 0u;         // the reference's (truncated) endpoint cell and its neighbours are all outside the grid
    if (gxc < 0 || gyc < 0)
        return 1u;         // truncation toward zero differs from the floor here: never certified (tab_eval, EDGE 2)
    if (a.lf_cells) {
        // likelihood field: every class is uniform (the ray scores the field value of its endpoint cell);
        // class 0 = the field is 0 on the cell and on its eight neighbours (certain whatever the exact cell)
        auto lfc = [&](int cx, int cy) -> int {
            return ((unsigned)cx < (unsigned)W && (unsigned)cy < (unsigned)H) ? (int)__ldg(a.lf_cells + (size_t)cy * gp + cx) : 0;
        };
        const int u = lfc(gxc, gyc);
        if (u > 0) return 1u + (unsigned)u;
        int mx = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) mx = max(mx, lfc(gxc + dx, gyc + dy));
        return mx > 0 ? 1u : 0u;
    }
    int f;
    if (gxc < W && gyc < H) {
        f = (int)__ldg(a.fast_cells + (size_t)gyc * gp + gxc);
    } else {        // beyond the high edges: the cell reads 0; is anything positive within two cells?
        bool any = false;
        for (int dy = -2; dy <= 2; ++dy)
            for (int dx = -2; dx <= 2; ++dx) any = any || rawc(gxc + dx, gyc + dy) > 0;
        f = any ? 0 : -1;
    }
    if (f < 0) return 0u;
    if (f > 0) return 1u + (unsigned)f;
    // sector s steps (ux, uy): 0 (+1,0) 1 (+1,+1) 2 (0,+1) 3 (-1,+1) 4 (-1,0) 5 (-1,-1) 6 (0,-1) 7 (+1,-1)
    n[0] = rawc(gxc + 1, gyc);     n[1] = rawc(gxc + 1, gyc + 1); n[2] = rawc(gxc, gyc + 1);
    n[3] = rawc(gxc - 1, gyc + 1); n[4] = rawc(gxc - 1, gyc);     n[5] = rawc(gxc - 1, gyc - 1);
    n[6] = rawc(gxc, gyc - 1);     n[7] = rawc(gxc + 1, gyc - 1);
    int mx = 0;
#pragma unroll
    for (int sct = 0; sct < 8; ++sct) mx = max(mx, n[sct]);
    return mx <= 0 ? 1u : kTabNeedsEntry;
}

// The score-table entry of a cell from its neighbours: per sector o1 > 0 ? o1 : max(o2, 0) (sensor_model.cpp:48-57).
__device__ __forceinline__ unsigned long long tab_entry(const int n[8])
{
    unsigned long long pack = 0ull;
#pragma unroll
    for (int sct = 0; sct < 8; ++sct) {
        const int o1 = n[(sct + 4) & 7], o2 = n[sct];
        const int v = o1 > 0 ? o1 : max(o2, 0);
        pack |= (unsigned long long)v << (8 * sct);
    }
    return pack;
}

// The CLASS MAP: tab_classify of every cell of the grid plus a four-cell apron (cell (gx, gy) at
// cls[(gy + 4) * cpitch + gx + 4]; 255 = the cell gets a score-table entry, which is pack[same index]), kept current with
// the mirror (refresh_fast_map) so that building a window is a copy plus the numbering of its entries.
__global__ void derive_class_map_kernel(const TabArgs a, uint8_t* cls, unsigned long long* pack, int cpitch, int x0, int y0,
                                        int w, int h)
{
    const int total = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int gyc = y0 + i / w, gxc = x0 + i % w;
        int n[8];
        const unsigned K = tab_classify(a, gxc, gyc, n);
        const size_t at = (size_t)(gyc + kTabApron) * cpitch + (gxc + kTabApron);
        cls[at] = (uint8_t)(K == kTabNeedsEntry ? 255u : K);
        if (K == kTabNeedsEntry) pack[at] = tab_entry(n);
    }
}

// Builds the class tile K and the score table T of the window in pl (see the head of this file) from the class map.
// All threads of the CTA; synchronises inside; the caller synchronises after it.  s_count: entries of T in use beyond
// the fixed ones, s_overflow: T is full (the caller then scores the window's particles exactly).
//   One warp per row, 64-cell segment by segment (four segments' loads in flight): the lanes fetch the classes of
//   their two cells, ballots number the cells that get an entry and lane 0 reserves the segment's run of entries; an
//   entry remembers its cell.  WIDE: the tile keeps 129 + the number within the segment and the run's base goes to
//   the row's segment bases; else the tile keeps the entry's index.  Then the entries fetch their eight scores from
//   the class map's pack array (independent loads).
template <bool WIDE>
__device__ __forceinline__ void tab_build_window(const TabPlan& pl, const TabArgs& a, uint16_t* ktile,
                                                 unsigned long long* ttab, unsigned k0, int* s_count, int* s_overflow)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* k8 = reinterpret_cast<uint8_t*>(ktile);
    for (int i = tid; i < kTabFixed; i += kTabThreads)
        ttab[i] = i >= 2 ? 0x0101010101010101ull * (unsigned long long)(2 * (i - 1)) : 0ull;
    const int w = pl.w, h = pl.h, pitch = pl.pitch_k, cap = pl.cap_entries, nseg = pl.nseg, seg_off = pl.seg_off;
    const int x0 = pl.x0, y0 = pl.y0, cpitch = a.cpitch;
    const int bx = (int)(pl.bias_x & 63u);                       // segment s starts at tile column 64 s - bx
    const int wlim = min(w, a.grid.width + kTabApron - x0);      // (a batch window wider than the grid + apron)
    const int hlim = a.grid.height + kTabApron - y0;
    const uint8_t* src = a.cls + (size_t)(y0 + kTabApron) * cpitch + (x0 + kTabApron);
    const unsigned below = (1u << lane) - 1u;
    for (int ty = warp; ty < h; ty += kTabWarps) {
        const uint8_t* row = src + (size_t)ty * cpitch;
        const bool row_in = ty < hlim;
        const unsigned at_row = (unsigned)(y0 + ty + kTabApron) * (unsigned)cpitch + (unsigned)(x0 + kTabApron);
        for (int sg0 = 0; sg0 < nseg; sg0 += 4) {
            unsigned c[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int tx = (sg0 + (j >> 1)) * 64 - bx + (j & 1) * 32 + lane;
                c[j] = (row_in && tx >= 0 && tx < wlim) ? (unsigned)__ldg(row + tx) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sg = sg0 + j;
                if (sg >= nseg) break;
                const int tx0 = sg * 64 - bx + lane, tx1 = tx0 + 32;
                const bool n0 = c[2 * j] == 255u, n1 = c[2 * j + 1] == 255u;
                const unsigned need0 = __ballot_sync(0xffffffffu, n0), need1 = __ballot_sync(0xffffffffu, n1);
                const int c0 = __popc(need0), cnt = c0 + __popc(need1);
                int first = 0;
                bool room = true;
                if (cnt > 0) {
                    if (lane == 0) first = kTabFixed + atomicAdd(s_count, cnt);
                    first = __shfl_sync(0xffffffffu, first, 0);
                    room = first + cnt <= cap;
                    if (!room && lane == 0) *s_overflow = 1;
                }
                if (WIDE && lane == 0)
                    *reinterpret_cast<uint16_t*>(k8 + ty * pitch + seg_off + 2 * sg) = (uint16_t)(k0 + (unsigned)first - (unsigned)kTabFixed);
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const int tx = hlf ? tx1 : tx0;
                    if (tx < 0 || tx >= w) continue;
                    unsigned K = c[2 * j + hlf];
                    if (K == 255u) {
                        const int id = hlf ? c0 + __popc(need1 & below) : __popc(need0 & below);
                        if (room) {
                            ttab[first + id] = (unsigned long long)(at_row + (unsigned)tx);
                            K = WIDE ? (unsigned)(kTabFixed + id) : (unsigned)(first + id);
                        } else {
                            K = 1u;
                        }
                    }
                    if (WIDE) k8[ty * pitch + tx] = (uint8_t)K;
                    else ktile[ty * pitch + tx] = (uint16_t)(K + k0);
                }
            }
        }
    }
    __syncthreads();
    // (not after an overflow: the caller scores the window exactly, and the entries past the last complete run do not
    // know their cells)
    const int used = *s_overflow ? 0 : kTabFixed + *s_count;
    for (int e = kTabFixed + tid; e < used; e += kTabThreads) {
        const unsigned at = (unsigned)ttab[e];
        ttab[e] = __ldg(a.pack + at);
    }
}

// BATCH = false: one window for the whole slice; units of 32 particles are dealt over all warps of the grid.
// BATCH = true:  CTAs take batches of a.batch consecutive particles from a counter (a.build_info[2]), place the
//                plan's fixed-size window for each (tab_batch_window) and rebuild K and T; the batch's units are dealt
//                over the CTA's warps.  A build is w h cell classifications for a.batch x beams evaluations (config 5:
//                63 K cells for 1.46 M evaluations).
// WIDE:          8-bit class tile + segment bases (tab_eval), for windows whose 16-bit tile does not fit.
template <bool INTERP, bool COUNT, bool BATCH, bool WIDE>
__global__ void __launch_bounds__(kTabThreads, 1) score_table_kernel(const TabArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TabPlan s_plan;
    __shared__ TabCold s_cold;
    __shared__ int s_count, s_overflow;
    __shared__ long long s_batch;
    __shared__ int s_batch_ok, s_nb;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb_all = a.num_beams;
    const int nb4 = (nb_all + 3) & ~3;
    TabBeam* sfast = reinterpret_cast<TabBeam*>(smem);
    Beam* sbeams = reinterpret_cast<Beam*>(smem + (size_t)nb4 * sizeof(TabBeam));
    float* sd8 = reinterpret_cast<float*>(smem + (size_t)nb4 * (sizeof(TabBeam) + sizeof(Beam)));
    uint16_t* squeue = reinterpret_cast<uint16_t*>(sd8 + nb4);
    int* swacc = reinterpret_cast<int*>(squeue + kTabWarps * kTabQueue);

    if (tid == 0) {
        s_plan = *a.plan;
        s_count = 0; s_overflow = 0;
    }
    swacc[tid] = 0;
    __syncthreads();
    const TabPlan& pl = s_plan;
    const bool table = pl.ok != 0 && (pl.batch != 0) == BATCH && (pl.wide != 0) == WIDE;
    const float cpm = a.grid.cells_per_meter;
    // The beams the table pass evaluates, compacted: those table_cull_kernel found to score 0 for every particle of the
    // slice are left out (their position in the compacted list is kept in the queue words for a moment).
    const bool culling = !BATCH && table && pl.cull_ok && a.cull != nullptr;
    if (warp == 0) {
        int base = 0;
        for (int i0 = 0; i0 < nb_all; i0 += 32) {
            const int i = i0 + lane;
            const bool keep = i < nb_all && !(culling && a.cull[i]);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (i < nb_all) squeue[i] = (uint16_t)(keep ? base + __popc(m & ((1u << lane) - 1u)) : 0xffffu);
            base += __popc(m);
        }
        if (lane == 0) s_nb = base;
    }
    __syncthreads();
    const int nb = s_nb;
    uint16_t my_slot[2];                                     // (kTabMaxBeams <= 2 * kTabThreads)
    for (int k = 0; k < 2; ++k) my_slot[k] = tid + k * kTabThreads < nb_all ? squeue[tid + k * kTabThreads] : (uint16_t)0xffffu;
    __syncthreads();                                         // the queue words are free again
    for (int k = 0; k < 2; ++k) {
        if (my_slot[k] == 0xffffu) continue;
        const int i = my_slot[k];
        const Beam b = a.beams[tid + k * kTabThreads];
        sbeams[i] = b;
        TabBeam f;
        const float rc = __fmul_rn(b.range, cpm);
        // the table pass folds the beam angle into [-pi, pi] (the same direction): its SFU arguments then stay within
        // +-(2 pi + 0.2) whatever the particle's heading (lidar angles run to 2 pi; kFastTrigErr is measured to 9.5)
        const double thd = (double)b.theta;
        f.ratio = (float)b.ratio;
        f.theta = (float)(thd > 3.14159265358979323846 ? thd - 6.28318530717958647692
                                                       : (thd < -3.14159265358979323846 ? thd + 6.28318530717958647692 : thd));
        f.rcx = (float)((double)rc * pl.inv_sx); f.rcy = (float)((double)rc * pl.inv_sy);
        // |2|px| - |py|| = sqrt5 rc |sin(angle to the octant boundary)| must exceed T3: angular band, in u8 units
        const double xq = (double)pl.t3 / (2.2360679 * (double)rc * (1.0 - 1e-6));
        const double d8 = (xq >= 0.3 || !(xq >= 0.0)) ? 4.0 : (asin(xq) + 2e-5) * 1.2732395447351628 * (1.0 + 1e-6);
        sd8[i] = (float)d8 + 1e-6f;
        sfast[i] = f;
    }

    uint16_t* ktile = reinterpret_cast<uint16_t*>(smem + pl.off_k);
    unsigned long long* ttab = reinterpret_cast<unsigned long long*>(smem + pl.off_t);
    const unsigned k0 = (unsigned)__cvta_generic_to_shared(ttab) >> 3;

    TabConst tc;
    tc.mul_x = pl.mul_x; tc.mul_y = pl.mul_y; tc.frac_thr = pl.frac_thr;
    tc.pitch2 = (unsigned)pl.pitch_k * (WIDE ? 1u : 2u);
    const unsigned sk = (unsigned)__cvta_generic_to_shared(ktile);
    tc.kbase = sk - pl.bias_y * tc.pitch2 - pl.bias_x * (WIDE ? 1u : 2u);
    tc.sdelta = pl.bias_x + (unsigned)pl.seg_off - (pl.bias_x >> 6) * 2u;
    tc.k0 = k0;
    tc.k129 = k0 + (unsigned)kTabFixed;
    tc.lf = a.lf_cells ? 1u : 0u;
    // (an IMAD takes one uniform-register operand: keep the row pitch in a vector register, or every cell address of
    // the loop pays a move)
    {
        unsigned z;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(z));
        tc.pitch2 += z >> 5;            // + 0, but no longer provably warp-uniform
    }
    const double gx = (double)a.grid.origin_x, gy = (double)a.grid.origin_y, cpm_d = (double)cpm;
    const int nwords = (nb + 31) / 32;
    uint16_t* q = squeue + warp * kTabQueue;
    int* wacc = swacc + warp * 32;
    int gathers = 0;

    // One unit = the 32 particles [p0, p0 + 32) below p_end, beams of the 32-beam words [w0, w1); add_mode: the unit is
    // one of several beam ranges of the same particles, whose partial sums add up in score2 (zeroed by the host).
    // (particles are addressed relative to the slice, in 32 bits: the loop is short of registers)
    auto score_unit = [&](unsigned i0, unsigned i_end, int w0, int w1, bool add_mode, bool degrade, bool first_part) {
        const bool live = i0 + lane < i_end;
        float xa = 0.f, ya = 0.f, tha = 0.f, xb = 0.f, yb = 0.f, thb = 0.f;
        if (live) {
            const long long p = a.lo + (long long)(i0 + lane);
            xa = a.x[p]; ya = a.y[p]; tha = a.th[p]; xb = a.px[p]; yb = a.py[p]; thb = a.pth[p];
        }
        TabBase fb = make_tab_base<INTERP>(xa, ya, tha, xb, yb, thb, gx, gy, cpm_d, pl);
        {
            const unsigned bad = __ballot_sync(0xffffffffu, live && !fb.ok);
            if (bad && lane == 0 && first_part) atomicAdd(a.build_info + 4, __popc(bad));
        }
        fb.ok = fb.ok && live && !degrade;
        const int edge = __reduce_max_sync(0xffffffffu, fb.ok ? fb.edge : 0);
        int acc = 0, qn = 0;
        unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;      // uncertain-beam bits of four consecutive 32-beam words
        for (int w = w0; w < w1; ++w) {
            uint32_t m = 0;
            const int kend = min(32, nb - w * 32);
            if (fb.ok) {
                auto run = [&](auto edge_tag) {
                    constexpr int EDGE = decltype(edge_tag)::value;
                    const TabBeam* bw = sfast + w * 32;
                    const float* dw = sd8 + w * 32;
                    uint32_t bit = 1u;
MCL_UNROLL(MCL_TAB_UNROLL)
                    for (int k = 0; k < kend; ++k) {
                        int add;
                        const bool certain = tab_eval<INTERP, COUNT, EDGE, WIDE>(fb, bw[k], dw[k], tc, pl, add, gathers);
                        acc += add;
                        m |= certain ? 0u : bit;
                        bit += bit;
                    }
                };
                if (edge == 0) run(std::integral_constant<int, 0>{});
                else if (edge == 1) run(std::integral_constant<int, 1>{});
                else run(std::integral_constant<int, 2>{});
            } else if (live) {
                m = kend == 32 ? 0xffffffffu : ((1u << kend) - 1u);
            }
            const int part = (w - w0) & 3;
            if (part == 0) m0 = m; else if (part == 1) m1 = m; else if (part == 2) m2 = m; else m3 = m;
            if (part == 3 || w == w1 - 1) {
                // hand the uncertain beams of these (up to) four words to the exact path
                if (__any_sync(0xffffffffu, (m0 | m1 | m2 | m3) != 0u))
                    qn = tab_defer_words<INTERP, COUNT>(m0, m1, m2, m3, w - part, qn, xa, ya, tha, xb, yb, thb, &s_cold, q, wacc);
                m0 = m1 = m2 = m3 = 0u;
            }
        }
        if (qn > 0)
            tab_drain_round<INTERP, COUNT>(lane < qn ? q[lane] : 0u, lane < qn, xa, ya, tha, xb, yb, thb, &s_cold, wacc);
        __syncwarp();
        acc += wacc[lane];
        wacc[lane] = 0;
        if (COUNT && live && first_part) gathers += 3 * (nb_all - nb);      // a culled beam reads three cells in the reference
        if (live) {
            const long long p = a.lo + (long long)(i0 + lane);
            if (add_mode) atomicAdd(a.score2 + p, acc);
            else a.score2[p] = acc;
        }
        __syncwarp();
    };

    if (tid == 0) {
        s_cold.grid = a.grid; s_cold.sbeams = sbeams; s_cold.deferred = 0ull; s_cold.gathers = 0ull;
        s_cold.lf = a.lf_cells ? 1 : 0;
        if (a.lf_cells) s_cold.grid.cells = a.lf_cells;         // the exact path reads the field outside the window
        s_cold.ktile = nullptr;
        s_cold.w = pl.w; s_cold.h = pl.h; s_cold.pitch_k = pl.pitch_k;
        s_cold.k0 = k0;
        s_cold.wide = WIDE ? 1 : 0;
    }

    if (!BATCH) {
        // ---- K tile and score table of the slice's window --------------------------------------------------------
        if (table) tab_build_window<WIDE>(pl, a, ktile, ttab, k0, &s_count, &s_overflow);
        __syncthreads();
        const bool degrade = !table || s_overflow != 0;
        if (tid == 0) {
            if (table) {
                atomicMax(a.build_info + 0, kTabFixed + s_count);
                if (s_overflow) atomicAdd(a.build_info + 1, 1);
            }
            s_cold.ktile = degrade ? nullptr : ktile;
            s_cold.x0 = pl.x0; s_cold.y0 = pl.y0;
        }
        __syncthreads();

        // work unit = 32 consecutive particles = one warp; units are dealt round-robin over CTAs, then over warps.  The
        // last, partial round of units would leave most warps idle while a few finish a whole unit (with 2 M particles
        // per GPU that is 6 % of the kernel): those units are split by beams into up to S parts, one per otherwise idle
        // warp, which add their partial sums to score2 (zeroed by the host for exactly those particles:
        // table_tail_split).
        const long long nunits = (a.hi - a.lo + 31) / 32;
        long long whole_units, items;
        int split;
        // (the split is the host's: it depends on the scan's beam count, not on how many beams survive the culling)
        table_tail_split(nunits, (long long)gridDim.x * kTabWarps, (nb_all + 31) / 32, whole_units, split, items);
        for (long long item = (long long)blockIdx.x + (long long)gridDim.x * warp; item < items;
             item += (long long)gridDim.x * kTabWarps) {
            long long unit = item;
            int w0 = 0, w1 = nwords, part = 0;
            if (item >= whole_units) {
                const long long sub = item - whole_units;
                unit = whole_units + sub / split;
                part = (int)(sub - (sub / split) * split);
                w0 = part * nwords / split; w1 = (part + 1) * nwords / split;
            }
            score_unit((unsigned)(unit * 32), (unsigned)(a.hi - a.lo), w0, w1, item >= whole_units, degrade, part == 0);
        }
    } else {
        const long long nbatches = (a.hi - a.lo + a.batch - 1) / a.batch;
        int max_count = 0, overflows = 0;
        for (;;) {
            __syncthreads();                    // everyone is done with the previous batch's window (and s_batch)
            if (tid == 0) {
                const long long b = (long long)atomicAdd(a.build_info + 2, 1);
                s_batch = b;
                s_count = 0; s_overflow = 0;
                s_batch_ok = (table && b < nbatches && tab_batch_window(s_plan, a.bboxes[b], a.grid)) ? 1 : 0;
            }
            __syncthreads();
            const long long b = s_batch;
            if (b >= nbatches) break;
            const bool have = s_batch_ok != 0;
            if (have) tab_build_window<WIDE>(pl, a, ktile, ttab, k0, &s_count, &s_overflow);
            __syncthreads();
            const bool degrade = !have || s_overflow != 0;
            if (tid == 0) {
                if (!have) atomicAdd(a.build_info + 3, 1);
                if (have) { max_count = max(max_count, kTabFixed + s_count); overflows += s_overflow; }
                s_cold.ktile = degrade ? nullptr : ktile;
                s_cold.x0 = pl.x0; s_cold.y0 = pl.y0;
            }
            __syncthreads();
            const unsigned first = (unsigned)(b * a.batch);
            const unsigned last = (unsigned)((b + 1) * a.batch < a.hi - a.lo ? (b + 1) * a.batch : a.hi - a.lo);
            for (unsigned i0 = first + warp * 32; i0 < last; i0 += kTabThreads)
                score_unit(i0, last, 0, nwords, false, degrade, true);
        }
        if (tid == 0 && table) {
            atomicMax(a.build_info + 0, max_count);
            if (overflows) atomicAdd(a.build_info + 1, overflows);
        }
    }
    __syncthreads();
    if (COUNT) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gathers += __shfl_xor_sync(0xffffffffu, gathers, off);
        if (lane == 0) atomicAdd(a.gather_counter, (unsigned long long)gathers);
        if (tid == 0) atomicAdd(a.gather_counter, s_cold.gathers);
    }
    if (tid == 0 && s_cold.deferred) atomicAdd(a.deferred_counter, s_cold.deferred);
}

__host__ __device__ inline size_t table_fixed_smem(int num_beams)
{
    size_t b = (size_t)((num_beams + 3) & ~3) * (sizeof(TabBeam) + sizeof(Beam) + sizeof(float)) +
               (size_t)kTabWarps * kTabQueue * sizeof(uint16_t) + (size_t)kTabThreads * sizeof(int);
    return (b + 15) & ~(size_t)15;
}

}  // namespace mcl
