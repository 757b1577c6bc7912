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
    int pitch_k;                 // K entries per row (2 bytes each); pitch_k / 2 is odd (rows spread over the banks)
    int cap_entries;             // T capacity, the fixed ones included
    unsigned off_k, off_t;       // byte offsets of K and T in dynamic shared memory
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
};

struct TabPlanIn {
    DevGrid grid;
    float max_range, min_range, max_abs_theta;
    double ratio_lo, ratio_hi;
    int num_beams;
    int scan_finite;
    int allow;
    int smem_total, smem_fixed;
};

// One thread: box -> window -> budget.  Resets the box for the next bbox_kernel.
__global__ void table_plan_kernel(const TabPlanIn in, int* box, TabPlan* out)
{
    TabPlan pl;
    memset(&pl, 0, sizeof(pl));
    auto unorder = [](int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); };
    const float mnx = unorder(box[0]), mny = unorder(box[1]), mxx = unorder(box[2]), mxy = unorder(box[3]);
    box[0] = 0x7fffffff; box[1] = 0x7fffffff; box[2] = (int)0x80000000; box[3] = (int)0x80000000;
    const double cpm = (double)in.grid.cells_per_meter;
    const double Rc = (double)in.max_range * cpm;
    const double rho_max = fmax(fabs(in.ratio_lo), fabs(in.ratio_hi));
    do {
        if (!in.allow) { pl.reason = 5; break; }
        if (!in.scan_finite || in.num_beams < 1 || in.num_beams > kTabMaxBeams || !isfinite(Rc) || !(cpm > 0.0) ||
            !((double)in.min_range * cpm >= 2.5) || !(in.max_abs_theta <= 6.3f) || !(in.ratio_lo >= -1.0) ||
            !(in.ratio_hi <= 2.0)) { pl.reason = 1; break; }
        if (!(isfinite(mnx) && isfinite(mny) && isfinite(mxx) && isfinite(mxy)) || mnx > mxx || mny > mxy) { pl.reason = 2; break; }
        // window = bounding box of poses and parents +- (longest ray + 6 cells), not clipped to the grid (cells outside
        // read 0).  +6: every particle of the cloud then passes make_tab_base's interior test, whose reach carries
        // 4 cells of margin plus 1.5 for the window's border.
        const double reach = Rc + 6.0;
        const double cx0 = floor(((double)mnx - (double)in.grid.origin_x) * cpm - reach);
        const double cy0 = floor(((double)mny - (double)in.grid.origin_y) * cpm - reach);
        const double cx1 = ceil(((double)mxx - (double)in.grid.origin_x) * cpm + reach);
        const double cy1 = ceil(((double)mxy - (double)in.grid.origin_y) * cpm + reach);
        if (!(fabs(cx0) < 1.0e6 && fabs(cy0) < 1.0e6 && cx1 - cx0 < 8192.0 && cy1 - cy0 < 8192.0)) { pl.reason = 3; break; }
        const long long ux0 = (long long)cx0, uy0 = (long long)cy0, ux1 = (long long)cx1, uy1 = (long long)cy1;
        // clipped to the grid plus four cells: everything beyond is class 0, and so are the clipped window's border cells,
        // onto which the saturating coordinate arithmetic maps rays that leave it
        const long long x0 = ux0 > -4 ? ux0 : -4, y0 = uy0 > -4 ? uy0 : -4;
        const long long x1 = ux1 < in.grid.width + 3 ? ux1 : in.grid.width + 3, y1 = uy1 < in.grid.height + 3 ? uy1 : in.grid.height + 3;
        const long long tw = x1 - x0 + 1, th = y1 - y0 + 1;
        if (tw < 8 || th < 8) { pl.reason = 3; break; }       // (a cloud whose rays cannot reach the grid)
        long long pitch_k = (tw + 1) & ~1ll;
        if (((pitch_k >> 1) & 1) == 0) pitch_k += 2;
        const long long k_bytes = (pitch_k * 2 * th + 15) & ~15ll;
        const long long room = (long long)in.smem_total - in.smem_fixed - k_bytes - 64;
        pl.need_bytes = (int)fmin(2.0e9, (double)(in.smem_fixed + k_bytes + 64 + 8 * (kTabFixed + 256)));
        if (room < 8 * (kTabFixed + 256)) { pl.reason = 3; break; }
        // error budget (cells).  Reference vs the real-valued model: as fast_plan (mcl_engine.cu).  Float model vs the
        // same: roundings of dS, rho and rc, the angle roundings and the measured SFU error; its coordinate roundings
        // (the normalised robot coordinate, the ratio FFMA, the endpoint FFMA, the +1.0) are below 2^-24 of the
        // window's extent each.
        const double u = 5.9604644775390625e-08;
        const double Cm = (double)(ux1 > uy1 ? ux1 : uy1) + 2.0;
        if (Cm > 16000.0) { pl.reason = 3; break; }
        const double Xm = Cm / cpm + fmax(fabs((double)in.grid.origin_x), fabs((double)in.grid.origin_y));
        const double max_shift = 64.0;
        const double Ce = Cm + Rc;
        const double e_ref = cpm * u * Xm + 2.0 * u * Ce + 2.0 * u * Rc + Rc * (20.0 * u + 1.2e-7) + 1e-9;
        const double e_apx = 6.0 * u * (double)(tw > th ? tw : th) + (1.0 + 2.0 * rho_max) * u * max_shift + 3.0 * u * Rc +
                             Rc * ((3.14159265358979 * (3.0 * rho_max + 1.0) + 9.5) * u + (double)kFastTrigErr);
        const double eps = 1.25 * (e_ref + e_apx) + 1e-6;
        const double kappa = eps + 1e-5;
        if (kappa > 1.0 / 16.0) { pl.reason = 4; break; }
        pl.ok = 1;
        pl.x0 = (int)x0; pl.y0 = (int)y0; pl.w = (int)tw; pl.h = (int)th; pl.pitch_k = (int)pitch_k;
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
        pl.ang_room = 9.5f - in.max_abs_theta;             // kFastTrigErr is measured for |angle| <= 9.5
        pl.ulo_x = (float)((double)ux0 + 1.5 + (double)pl.reach); pl.uhi_x = (float)((double)(ux1 + 1) - 1.5 - (double)pl.reach);
        pl.ulo_y = (float)((double)uy0 + 1.5 + (double)pl.reach); pl.uhi_y = (float)((double)(uy1 + 1) - 1.5 - (double)pl.reach);
        pl.hmin_x = (unsigned)((long long)pl.bias_x - x0);
        pl.hmin_y = (unsigned)((long long)pl.bias_y - y0);
        const double x2_min = 3.0 * eps + 2.0 * kappa + 1e-3;         // global coordinate the doubled endpoint must exceed
        pl.x2_lo_x = (float)((x2_min - (double)x0 + pl.off) * pl.inv_sx * (1.0 + 1e-6) + 1e-7);
        pl.x2_lo_y = (float)((x2_min - (double)y0 + pl.off) * pl.inv_sy * (1.0 + 1e-6) + 1e-7);
    } while (false);
    *out = pl;
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
template <bool INTERP, bool COUNT, int EDGE>
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
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(kaddr) : "r"(hy), "r"(tc.pitch2), "r"(tc.kbase));
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(kaddr) : "r"(hx), "r"(kaddr));
    bool x2_pos = true;
    if (EDGE >= 1)      // the doubled endpoint must not have a negative coordinate (the reference truncates toward zero there)
        x2_pos = (__fmaf_rn(b.rcx, c, nx) >= pl.x2_lo_x) & (__fmaf_rn(b.rcy, s, ny) >= pl.x2_lo_y);
    unsigned K;
    asm("ld.shared.u16 %0, [%1];" : "=r"(K) : "r"(kaddr));
    // sector: u8 = 8t - 2 round(4t) in [-1, 1] (t = a / 2pi) is the offset from the nearest axis, +-1 = 45 degrees;
    // n = 2 round(4t) + round(u8 * S) with S = 0.5 / B2, so the rounding flips exactly at the octant boundaries +-B2
    const float r8 = __fmaf_rn(a, kTabC8, 25165824.0f);                 // 1.5 * 2^24: ulp 2
    const float u8 = __fmaf_rn(a, kTabC8, -__fsub_rn(r8, 25165824.0f));
    const float wq = __fmaf_rn(u8, kTabS, __fsub_rn(r8, 12582912.0f));   // 1.5 * 2^23 + 2 round(4t): ulp 1
    // K carries the table's base (shared address / 8), so K * 8 + sector is the address of the score byte
    unsigned v, taddr;
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(taddr) : "r"(K), "r"(__float_as_uint(wq) & 7u));
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(taddr));
    const bool dir_ok = x2_pos & (fabsf(__fsub_rn(fabsf(u8), kTabB2)) > d8);
    const bool certain = (K == tc.k0) | (cell_ok & ((K < tc.k129) | dir_ok));
    add = certain ? (int)v : 0;
    if (COUNT) gathers += certain ? ((K - tc.k0 - 2u < 127u) | (tc.lf != 0u) ? 1 : 3) : 0;
    return certain;
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
    int* build_info;                // [0] table entries the window needs (max over CTAs), [1] CTAs that overflowed
    int num_peers;
    int32_t* peer_score[kMaxPeers];
};

// Cold-path state shared by the CTA (static shared memory): what the exact evaluations need.
struct TabCold {
    DevGrid grid;
    const Beam* sbeams;
    const uint16_t* ktile;      // class tile (null: not built -- read the mirror)
    int x0, y0, w, h, pitch_k;
    unsigned k0;
    int lf;                     // likelihood-field mode: a ray scores 2 u(endpoint cell), no neighbour steps
    unsigned long long deferred, gathers;
};

// Cell value as the score sees it: the log-odds where positive, else 0 (sensor_model.cpp:41-57 only ever tests
// "> 0").  Inside the window the class tile has it (classes 2..128 = occupied, log-odds class - 1); outside, the mirror.
__device__ __forceinline__ int tab_cell_value(const TabCold* c, int gx, int gy)
{
    const int tx = (int)((unsigned)gx - (unsigned)c->x0), ty = (int)((unsigned)gy - (unsigned)c->y0);
    if (c->ktile && (unsigned)tx < (unsigned)c->w && (unsigned)ty < (unsigned)c->h) {
        const unsigned cls = (unsigned)c->ktile[ty * c->pitch_k + tx] - c->k0;
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
// exact path.  Returns the new queue length (< 32).
template <bool INTERP, bool COUNT>
__device__ __noinline__ int tab_defer_words(unsigned m0, unsigned m1, unsigned m2, unsigned m3, int wbase, int qn,
                                            float xa, float ya, float tha, float xb, float yb, float thb, TabCold* cold,
                                            uint16_t* q, int* wacc)
{
    const int lane = threadIdx.x & 31;
    const int c = __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (qn + total <= kTabQueue) {
        int pos = qn + incl - c;
        unsigned m = m0;
        int beam0 = wbase * 32;
        for (int part = 0; part < 4; ++part) {
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                q[pos++] = (uint16_t)((lane << 11) | (beam0 + k));
            }
            m = part == 0 ? m1 : (part == 1 ? m2 : m3);
            beam0 += 32;
        }
        qn += total;
        __syncwarp();
        while (qn >= 32) {
            qn -= 32;
            tab_drain_round<INTERP, COUNT>(q[qn + lane], true, xa, ya, tha, xb, yb, thb, cold, wacc);
        }
        __syncwarp();
    } else {
        // burst (lanes outside the table pass's domain): every lane with bits left evaluates its own next beam; the
        // queue keeps what it held
        unsigned m = m0;
        int beam0 = wbase * 32;
        for (int part = 0; part < 4; ++part) {
            while (__any_sync(0xffffffffu, m != 0u)) {
                const bool has = m != 0u;
                const int k = has ? __ffs(m) - 1 : 0;
                if (has) m &= m - 1;
                tab_drain_round<INTERP, COUNT>((unsigned)((lane << 11) | (beam0 + k)), has, xa, ya, tha, xb, yb, thb, cold, wacc);
            }
            m = part == 0 ? m1 : (part == 1 ? m2 : m3);
            beam0 += 32;
        }
    }
    return qn;
}

#ifndef MCL_TAB_UNROLL
#define MCL_TAB_UNROLL 4
#endif

template <bool INTERP, bool COUNT>
__global__ void __launch_bounds__(kTabThreads, 1) score_table_kernel(const TabArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ TabPlan s_plan;
    __shared__ TabCold s_cold;
    __shared__ int s_count, s_overflow;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = a.num_beams;
    const int nb4 = (nb + 3) & ~3;
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
    const bool table = pl.ok != 0;
    const float cpm = a.grid.cells_per_meter;
    for (int i = tid; i < nb; i += kTabThreads) {
        const Beam b = a.beams[i];
        sbeams[i] = b;
        TabBeam f;
        const float rc = __fmul_rn(b.range, cpm);
        f.ratio = (float)b.ratio; f.theta = b.theta;
        f.rcx = (float)((double)rc * pl.inv_sx); f.rcy = (float)((double)rc * pl.inv_sy);
        // |2|px| - |py|| = sqrt5 rc |sin(angle to the octant boundary)| must exceed T3: angular band, in u8 units
        const double xq = (double)pl.t3 / (2.2360679 * (double)rc * (1.0 - 1e-6));
        const double d8 = (xq >= 0.3 || !(xq >= 0.0)) ? 4.0 : (asin(xq) + 2e-5) * 1.2732395447351628 * (1.0 + 1e-6);
        sd8[i] = (float)d8 + 1e-6f;
        sfast[i] = f;
    }

    // ---- K tile and score table of this window ------------------------------------------------------------------
    uint16_t* ktile = reinterpret_cast<uint16_t*>(smem + pl.off_k);
    unsigned long long* ttab = reinterpret_cast<unsigned long long*>(smem + pl.off_t);
    const unsigned k0 = (unsigned)__cvta_generic_to_shared(ttab) >> 3;
    if (table) {
        const int W = a.grid.width, H = a.grid.height, gp = a.grid.pitch;
        const int8_t* raw = a.grid.cells;
        auto rawc = [&](int gxc, int gyc) -> int {
            return ((unsigned)gxc < (unsigned)W && (unsigned)gyc < (unsigned)H) ? (int)__ldg(raw + (size_t)gyc * gp + gxc) : 0;
        };
        for (int i = tid; i < kTabFixed; i += kTabThreads)
            ttab[i] = i >= 2 ? 0x0101010101010101ull * (unsigned long long)(2 * (i - 1)) : 0ull;
        const int total = pl.w * pl.h;
        for (int i = tid; i < total; i += kTabThreads) {
            const int ty = i / pl.w, tx = i - ty * pl.w;
            const int gxc = pl.x0 + tx, gyc = pl.y0 + ty;
            unsigned K;
            if (gxc <= -4 || gyc <= -4 || gxc >= W + 3 || gyc >= H + 3) {
                K = 0u;        // the reference's (truncated) endpoint cell and its neighbours are all outside the grid
            } else if (gxc < 0 || gyc < 0) {
                K = 1u;        // truncation toward zero differs from the floor here: never certified (tab_eval, EDGE 2)
            } else if (a.lf_cells) {
                // likelihood field: every class is uniform (the ray scores the field value of its endpoint cell);
                // class 0 = the field is 0 on the cell and on its eight neighbours (certain whatever the exact cell)
                auto lfc = [&](int cx, int cy) -> int {
                    return ((unsigned)cx < (unsigned)W && (unsigned)cy < (unsigned)H) ? (int)__ldg(a.lf_cells + (size_t)cy * gp + cx) : 0;
                };
                const int u = lfc(gxc, gyc);
                if (u > 0) K = 1u + (unsigned)u;
                else {
                    int mx = 0;
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) mx = max(mx, lfc(gxc + dx, gyc + dy));
                    K = mx > 0 ? 1u : 0u;
                }
            } else {
                int f;
                if (gxc < W && gyc < H) {
                    f = (int)__ldg(a.fast_cells + (size_t)gyc * gp + gxc);
                } else {        // beyond the high edges: the cell reads 0; is anything positive within two cells?
                    bool any = false;
                    for (int dy = -2; dy <= 2; ++dy)
                        for (int dx = -2; dx <= 2; ++dx) any = any || rawc(gxc + dx, gyc + dy) > 0;
                    f = any ? 0 : -1;
                }
                if (f < 0) K = 0u;
                else if (f > 0) K = 1u + (unsigned)f;
                else {
                    // sector s steps (ux, uy): 0 (+1,0) 1 (+1,+1) 2 (0,+1) 3 (-1,+1) 4 (-1,0) 5 (-1,-1) 6 (0,-1) 7 (+1,-1)
                    int n[8];
                    n[0] = rawc(gxc + 1, gyc);     n[1] = rawc(gxc + 1, gyc + 1); n[2] = rawc(gxc, gyc + 1);
                    n[3] = rawc(gxc - 1, gyc + 1); n[4] = rawc(gxc - 1, gyc);     n[5] = rawc(gxc - 1, gyc - 1);
                    n[6] = rawc(gxc, gyc - 1);     n[7] = rawc(gxc + 1, gyc - 1);
                    int mx = 0;
#pragma unroll
                    for (int sct = 0; sct < 8; ++sct) mx = max(mx, n[sct]);
                    if (mx <= 0) K = 1u;
                    else {
                        const int e = kTabFixed + atomicAdd(&s_count, 1);
                        if (e < pl.cap_entries) {
                            unsigned long long pack = 0ull;
#pragma unroll
                            for (int sct = 0; sct < 8; ++sct) {
                                const int o1 = n[(sct + 4) & 7], o2 = n[sct];      // sensor_model.cpp:48-57
                                const int v = o1 > 0 ? o1 : max(o2, 0);
                                pack |= (unsigned long long)v << (8 * sct);
                            }
                            ttab[e] = pack;
                            K = (unsigned)e;
                        } else {
                            K = 1u;
                            s_overflow = 1;
                        }
                    }
                }
            }
            ktile[ty * pl.pitch_k + tx] = (uint16_t)(K + k0);
        }
    }
    __syncthreads();
    const bool degrade = !table || s_overflow != 0;
    if (tid == 0) {
        if (table) {
            atomicMax(a.build_info + 0, kTabFixed + s_count);
            if (s_overflow) atomicAdd(a.build_info + 1, 1);
        }
        s_cold.grid = a.grid; s_cold.sbeams = sbeams; s_cold.deferred = 0ull; s_cold.gathers = 0ull;
        s_cold.lf = a.lf_cells ? 1 : 0;
        if (a.lf_cells) s_cold.grid.cells = a.lf_cells;         // the exact path reads the field outside the window
        s_cold.ktile = degrade ? nullptr : ktile;
        s_cold.x0 = pl.x0; s_cold.y0 = pl.y0; s_cold.w = pl.w; s_cold.h = pl.h; s_cold.pitch_k = pl.pitch_k;
        s_cold.k0 = k0;
    }
    __syncthreads();

    TabConst tc;
    tc.mul_x = pl.mul_x; tc.mul_y = pl.mul_y; tc.frac_thr = pl.frac_thr;
    tc.pitch2 = (unsigned)pl.pitch_k * 2u;
    const unsigned sk = (unsigned)__cvta_generic_to_shared(ktile);
    tc.kbase = sk - pl.bias_y * tc.pitch2 - pl.bias_x * 2u;
    tc.k0 = k0;
    tc.k129 = k0 + (unsigned)kTabFixed;
    tc.lf = a.lf_cells ? 1u : 0u;
    const double gx = (double)a.grid.origin_x, gy = (double)a.grid.origin_y, cpm_d = (double)cpm;
    const int nwords = (nb + 31) / 32;
    uint16_t* q = squeue + warp * kTabQueue;
    int* wacc = swacc + warp * 32;
    int gathers = 0;

    // work unit = 32 consecutive particles = one warp; units are dealt round-robin over CTAs, then over warps
    const long long nunits = (a.hi - a.lo + 31) / 32;
    for (long long unit = (long long)blockIdx.x + (long long)gridDim.x * warp; unit < nunits;
         unit += (long long)gridDim.x * kTabWarps) {
        const long long p = a.lo + unit * 32 + lane;
        const bool live = p < a.hi;
        float xa = 0.f, ya = 0.f, tha = 0.f, xb = 0.f, yb = 0.f, thb = 0.f;
        if (live) { xa = a.x[p]; ya = a.y[p]; tha = a.th[p]; xb = a.px[p]; yb = a.py[p]; thb = a.pth[p]; }
        TabBase fb = make_tab_base<INTERP>(xa, ya, tha, xb, yb, thb, gx, gy, cpm_d, pl);
        fb.ok = fb.ok && live && !degrade;
        const int edge = __reduce_max_sync(0xffffffffu, fb.ok ? fb.edge : 0);
        int acc = 0, qn = 0;
        unsigned m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;      // uncertain-beam bits of four consecutive 32-beam words
        for (int w = 0; w < nwords; ++w) {
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
                        const bool certain = tab_eval<INTERP, COUNT, EDGE>(fb, bw[k], dw[k], tc, pl, add, gathers);
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
            const int part = w & 3;
            if (part == 0) m0 = m; else if (part == 1) m1 = m; else if (part == 2) m2 = m; else m3 = m;
            if (part == 3 || w == nwords - 1) {
                // hand the uncertain beams of these (up to) four words to the exact path
                if (__any_sync(0xffffffffu, (m0 | m1 | m2 | m3) != 0u))
                    qn = tab_defer_words<INTERP, COUNT>(m0, m1, m2, m3, w & ~3, qn, xa, ya, tha, xb, yb, thb, &s_cold, q, wacc);
                m0 = m1 = m2 = m3 = 0u;
            }
        }
        if (qn > 0)
            tab_drain_round<INTERP, COUNT>(lane < qn ? q[lane] : 0u, lane < qn, xa, ya, tha, xb, yb, thb, &s_cold, wacc);
        __syncwarp();
        acc += wacc[lane];
        wacc[lane] = 0;
        if (live) {
            if (a.num_peers > 0) {
                for (int r = 0; r < a.num_peers; ++r) a.peer_score[r][p] = acc;     // final: to every rank
            } else {
                a.score2[p] = acc;
            }
        }
        __syncwarp();
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
