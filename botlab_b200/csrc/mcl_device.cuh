// Device-side arithmetic of the MCL hot path, written so that every rounding matches the reference's x86-64 SSE2
// build (no FMA contraction, mixed float/double evaluation exactly as SURVEY.md Appendix A records it).
// Compile with -fmad=false; the only fused operations are the explicit __fma_rn calls inside glibc_sincosf.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "glibc_sincosf.h"

namespace mcl {

constexpr double kPi = 3.14159265358979323846;       // M_PI
constexpr double kTwoPi = 2.0 * kPi;                 // 2.0*M_PI, exact doubling
constexpr float kPiF = 3.14159274101257324f;         // (float)M_PI: the smallest float > M_PI

struct DevGrid {
    const int8_t* cells;   // row-major, pitch bytes per row
    int width, height, pitch;
    float origin_x, origin_y, cells_per_meter;
};

// One valid beam of the current scan (range > min_range), prepared on the host once per scan:
// ratio = (times[n] - t_begin) / (t_end - t_begin) in double (common/interpolation.hpp:36).
struct __align__(16) Beam {
    float range;
    float theta;
    double ratio;
};

// common/angle_functions.hpp:12-24.  For a float a: (double)a > M_PI  <=>  a >= (float)M_PI, because (float)M_PI is the
// first float above M_PI; likewise on the negative side.  The 2*pi step is a double add rounded back to float.
__device__ __forceinline__ float wrap_to_pi(float a)
{
    if (a <= -kPiF) {
        do { a = (float)__dadd_rn((double)a, kTwoPi); } while (a <= -kPiF);
    } else if (a >= kPiF) {
        do { a = (float)__dadd_rn((double)a, -kTwoPi); } while (a >= kPiF);
    }
    return a;
}

// common/angle_functions.hpp:78-87 / :128-138 share this fold.
#ifndef MCL_FOLD_NOINLINE
#define MCL_FOLD_NOINLINE 1
#endif
#if MCL_FOLD_NOINLINE
// the fold is rare (headings within |dth| of +-pi): out of line, so the hot loop pays one compare and one branch
// instead of five predicated-off issue slots
__device__ __noinline__ double fold_pi_rare(double v) { return __dadd_rn(v, (v > 0) ? -kTwoPi : kTwoPi); }
__device__ __forceinline__ double fold_pi(double v)
{
    if (__builtin_expect(fabs(v) > kPi, 0)) v = fold_pi_rare(v);
    return v;
}
#else
__device__ __forceinline__ double fold_pi(double v)
{
    if (__builtin_expect(fabs(v) > kPi, 0)) v = __dadd_rn(v, (v > 0) ? -kTwoPi : kTwoPi);
    return v;
}
#endif

// float -> int the way the x86-64 host does it (cvttss2si): truncate; out of range or NaN -> INT_MIN.
__device__ __forceinline__ int f2i_x86(float v)
{
    return (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
}

__device__ __forceinline__ int grid_read(const DevGrid& g, int x, int y)
{
    if ((unsigned)x < (unsigned)g.width && (unsigned)y < (unsigned)g.height)
        return (int)__ldg(g.cells + (size_t)y * g.pitch + x);
    return 0;   // occupancy_grid.cpp:65-70: outside the grid reads as 0
}

// slam/sensor_model.cpp:61-86: the cell one Bresenham step from (x1,y1) toward (x2,y2).  Wrapping differences and
// 64-bit compares make INT_MIN inputs well defined and equal to the oracle's double compare.
__device__ __forceinline__ void bresenham_step(int x1, int y1, int x2, int y2, int& xo, int& yo)
{
    int dx = (int)((unsigned)x2 - (unsigned)x1);
    int dy = (int)((unsigned)y2 - (unsigned)y1);
    dx = dx < 0 ? (int)(0u - (unsigned)dx) : dx;
    dy = dy < 0 ? (int)(0u - (unsigned)dy) : dy;
    const int sx = x1 < x2 ? 1 : -1;
    const int sy = y1 < y2 ? 1 : -1;
    const long long e2 = 2ll * (long long)(int)((unsigned)dx - (unsigned)dy);
    xo = (e2 >= -(long long)dy) ? (int)((unsigned)x1 + (unsigned)sx) : x1;
    yo = (e2 <= (long long)dx) ? (int)((unsigned)y1 + (unsigned)sy) : y1;
}

// Per-particle constants of the ray construction (common/interpolation.hpp:24-50).
struct RayBase {
    float xa, ya, tha;          // pose (scan end)
    double xb, yb, thb;         // parent pose widened
    double dx, dy, dth;         // (float)(xa-xb) widened; angle_diff(tha, thb)
};

__device__ __forceinline__ RayBase make_ray_base(float xa, float ya, float tha, float xb, float yb, float thb)
{
    RayBase r;
    r.xa = xa; r.ya = ya; r.tha = tha;
    r.xb = (double)xb; r.yb = (double)yb; r.thb = (double)thb;
    r.dx = (double)__fsub_rn(xa, xb);                       // float subtraction, then widened (interpolation.hpp:39)
    r.dy = (double)__fsub_rn(ya, yb);
    r.dth = fold_pi(__dsub_rn((double)tha, (double)thb));   // angle_diff in double (:41)
    return r;
}

// wrap_to_pi with a float-only fast path for the common case a in (-3*pi, -pi] (see glibc_sincosf.h: exact by the
// exhaustive sweep); everything else takes the reference's double loop.
__device__ __forceinline__ float wrap_to_pi_fast(float a)
{
    if (a <= -kPiF || a >= kPiF) {
        const float t = __fadd_rn(a, GS_TWO_PI_HI);
        if (a > -9.42477f && a < 0.0f && fabsf(t) >= 9.5367431640625e-07f) a = __fadd_rn(t, GS_TWO_PI_LO);
        else a = wrap_to_pi(a);
    }
    return a;
}

// Map window a CTA reads from: either the shared-memory tile or the whole global mirror.  Cells with
// 1 <= x-x0 < w-1 and 1 <= y-y0 < h-1 ("interior") have all eight neighbours inside the window, so the endpoint and
// its two Bresenham neighbours are read without further checks.
struct Window {
    const int8_t* base;
    int x0, y0, w, h, pitch;
};

// Hoisted per-CTA grid constants.
struct GridConst {
    double gx, gy, cpm_d;
    float cpm;
    GsConsts trig;     // sincosf constants held in registers
};

// One particle-beam evaluation: moving_laser_scan.cpp:26-33 + sensor_model.cpp:28-59.  Returns the ray score in
// HALF units (2*odds, or o1, or o2): exact integers.  SMEM selects the window's address space.
// Fast path: all coordinates are small (so float->int conversion and 32-bit differences cannot overflow) and the
// endpoint is interior to the window.  Anything else -- NaN/huge poses, endpoints at the window edge or off the map --
// takes slow_ray(), the literal restatement with x86 conversion semantics and bounds-checked global reads.
template <bool SMEM>
__device__ __forceinline__ int window_read(const Window& w, int idx)
{
    return SMEM ? (int)w.base[idx] : (int)__ldg(w.base + idx);
}

__device__ __forceinline__ int slow_ray(const DevGrid& g, float px, float py, float sx, float sy, int* gathers)
{
    const int ex = f2i_x86(__fadd_rn(px, sx));
    const int ey = f2i_x86(__fadd_rn(py, sy));
    const int xx = f2i_x86(__fadd_rn(__fmul_rn(2.0f, px), sx));
    const int xy = f2i_x86(__fadd_rn(__fmul_rn(2.0f, py), sy));
    const int odds = grid_read(g, ex, ey);
    *gathers += 1;
    if (odds > 0) return 2 * odds;
    int ax, ay, bx, by;
    bresenham_step(ex, ey, f2i_x86(sx), f2i_x86(sy), ax, ay);
    bresenham_step(ex, ey, xx, xy, bx, by);
    const int o1 = grid_read(g, ax, ay);
    const int o2 = grid_read(g, bx, by);
    *gathers += 2;
    return o1 > 0 ? o1 : (o2 > 0 ? o2 : 0);
}

// Offset (in window bytes) of the cell one Bresenham step from (x1,y1) toward (x2,y2); 32-bit arithmetic, valid while
// all coordinates are below 2^28 in magnitude (sensor_model.cpp:61-86: step x iff 2dx >= dy, step y iff dx <= 2dy).
__device__ __forceinline__ int step_offset(int x1, int y1, int x2, int y2, int pitch)
{
    const int ddx = x2 - x1, ddy = y2 - y1;
    const int dx = abs(ddx), dy = abs(ddy);
    const int sx = ddx > 0 ? 1 : -1;
    const int sy = ddy > 0 ? pitch : -pitch;
    return ((2 * dx >= dy) ? sx : 0) + ((dx <= 2 * dy) ? sy : 0);
}

#ifndef MCL_EAGER_NEIGHBOURS
#define MCL_EAGER_NEIGHBOURS 1
#endif
// Ray construction + endpoint of one evaluation, with the reference's exact rounding sequence: sx, sy = robot cell
// coordinate (grid_utils.hpp:50-55), px, py = (range*cos)*cpm, (range*sin)*cpm, e1 = p + s before truncation.
template <bool INTERP>
__device__ __forceinline__ void exact_endpoint(const RayBase& p, const Beam& b, const GridConst& gc, float& sx, float& sy,
                                               float& px, float& py, float& e1x, float& e1y)
{
    float ox, oy, thr;
    if (INTERP) {
        ox = (float)__dadd_rn(p.xb, __dmul_rn(p.dx, b.ratio));                 // interpolation.hpp:45
        oy = (float)__dadd_rn(p.yb, __dmul_rn(p.dy, b.ratio));                 // :46
        thr = (float)fold_pi(__dadd_rn(p.thb, __dmul_rn(p.dth, b.ratio)));     // :47 angle_sum
    } else {
        ox = p.xa; oy = p.ya; thr = p.tha;                                      // :29-34 equal-utime early-out
    }
    const float th = wrap_to_pi_fast(__fsub_rn(thr, b.theta));                 // moving_laser_scan.cpp:33
    // grid_utils.hpp:50-55: double math, stored into Point<float>
    sx = (float)__dmul_rn(__dsub_rn((double)ox, gc.gx), gc.cpm_d);
    sy = (float)__dmul_rn(__dsub_rn((double)oy, gc.gy), gc.cpm_d);
    float s, c;
    glibc_sincosf_regs(gc.trig, th, &s, &c);
    px = __fmul_rn(__fmul_rn(b.range, c), gc.cpm);                             // (range*cos)*cpm, float
    py = __fmul_rn(__fmul_rn(b.range, s), gc.cpm);
    e1x = __fadd_rn(px, sx);                                                   // sensor_model.cpp:34 (before truncation)
    e1y = __fadd_rn(py, sy);                                                   // :35
}

template <bool INTERP, bool SMEM, bool COUNT>
__device__ __forceinline__ int score_beam(const RayBase& p, const Beam& b, const GridConst& gc, const Window& win,
                                          const DevGrid& grid, int& gathers)
{
    float sx, sy, px, py, e1x, e1y;
    exact_endpoint<INTERP>(p, b, gc, sx, sy, px, py, e1x, e1y);
    // One guard for every conversion below: the robot cell and the endpoint must be non-negative and below 2^22.
    // As unsigned bit patterns that is a single compare (negative values, -0, NaN and infinities all exceed it).
    const unsigned gmax = max(max(__float_as_uint(sx), __float_as_uint(sy)), max(__float_as_uint(e1x), __float_as_uint(e1y)));
    const bool small = gmax < 0x4A800000u;
    // Truncation of a non-negative float below 2^23 without the conversion unit: a round-toward-zero add of 2^23 leaves
    // floor(v) in the mantissa, so the bit pattern is 0x4B000000 + (int)v.
    const int bex = __float_as_int(__fadd_rz(e1x, 8388608.0f));
    const int bey = __float_as_int(__fadd_rz(e1y, 8388608.0f));
    const int bsx = __float_as_int(__fadd_rz(sx, 8388608.0f));
    const int bsy = __float_as_int(__fadd_rz(sy, 8388608.0f));
    const int tx = bex - (0x4B000000 + win.x0), ty = bey - (0x4B000000 + win.y0);
    const bool fast = small && ((unsigned)(tx - 1) < (unsigned)(win.w - 2)) && ((unsigned)(ty - 1) < (unsigned)(win.h - 2));
    if (__builtin_expect(!fast, 0)) {
        const int ex = f2i_x86(e1x), ey = f2i_x86(e1y);
        // endpoint two or more cells outside the grid: it and both neighbours read 0 (occupancy_grid.cpp:65-70)
        // (holds for any robot cell / extended point, even INT_MIN ones: a Bresenham step moves at most one cell)
        if ((unsigned)(ex + 1) > (unsigned)(grid.width + 1) || (unsigned)(ey + 1) > (unsigned)(grid.height + 1)) {
            if (COUNT) gathers += 3;
            return 0;
        }
        return slow_ray(grid, px, py, sx, sy, &gathers);
    }
    // :37-38  ((2*range)*cos)*cpm == 2*px exactly (power-of-two scaling commutes with rounding); may be negative
    const int xx = __float2int_rz(__fadd_rn(__fmul_rn(2.0f, px), sx));
    const int xy = __float2int_rz(__fadd_rn(__fmul_rn(2.0f, py), sy));
    const int idx = ty * win.pitch + tx;
    const int odds = window_read<SMEM>(win, idx);                              // :41
    const int off1 = step_offset(bex, bey, bsx, bsy, win.pitch);               // :48 toward the robot (same bias on both)
    const int off2 = step_offset(bex - 0x4B000000, bey - 0x4B000000, xx, xy, win.pitch);   // :49 away from the robot
    int o1 = 0, o2 = 0;
#if MCL_EAGER_NEIGHBOURS
    // all three reads in flight at once (the neighbours are interior to the window, so the loads are always safe);
    // the reference's "only if odds <= 0" is restored by the select below
    if (SMEM) {
        o1 = window_read<SMEM>(win, idx + off1);
        o2 = window_read<SMEM>(win, idx + off2);
    } else
#endif
    if (odds <= 0) {
        o1 = window_read<SMEM>(win, idx + off1);
        o2 = window_read<SMEM>(win, idx + off2);
    }
    if (COUNT) gathers += odds > 0 ? 1 : 3;
    return odds > 0 ? 2 * odds : (o1 > 0 ? o1 : max(o2, 0));
}

// ---------------------------------------------------------------------------------------------------------------
// Certified fast evaluation (first pass of the sensor model).
//
// The literal restatement above costs ~150 issue slots per evaluation, most of them spent reproducing the reference's
// float/double rounding sequence.  Its OUTPUT, however, is only three cell reads, selected by (a) the endpoint cell
// (floor of two coordinates) and (b) the octant class of the ray direction (the two Bresenham steps).  Both are
// insensitive to rounding unless a coordinate lies within the accumulated rounding error of an integer, or the
// direction lies within that error of an octant boundary.  The fast pass evaluates the same real-valued model
//     E = S_b + dS*rho + range*cpm*(cos, sin)(th_b + dth*rho - theta_beam)
// in plain float arithmetic with the SFU sine/cosine, in ~60 issue slots, and CERTIFIES its result:
//   * |fast - reference| <= eps for every coordinate, eps derived on the host from the map size, the scan's maximum
//     range and the measured SFU error bound (mcl_engine.cu: fast_plan; DESIGN.md section 5 has the budget);
//   * an endpoint coordinate whose distance to the nearest integer is <= eps, or that is not interior to the window and
//     inside the grid, is NOT certain;
//   * the reference's step is "x iff 2|ddx| >= |ddy|, y iff |ddx| <= 2|ddy|" on differences of floored coordinates,
//     each within 1 + eps of the real difference p = range*cpm*(cos, sin); so the step is certain iff
//     |2|px| - |py|| and |2|py| - |px|| exceed 3(1 + eps), and then the step toward the robot is the exact opposite of
//     the step toward the extended point (which needs the extended point at non-negative coordinates, where the
//     reference's truncation equals floor).  When the endpoint cell itself is occupied the steps are never looked at.
// An evaluation that is not certain contributes nothing here and sets its bit in the lane's mask; the second pass
// (score_deferred_kernel) re-evaluates exactly those with the literal restatement (score_beam).  Results are
// therefore identical to the reference's whatever the inputs: eps only moves work between the passes.
struct __align__(16) FastBeam {
    float ratio;     // (float)Beam::ratio
    float theta;
    float rc;        // range * cells_per_meter
    float pad;
};

struct FastPlan {
    int enabled;
    // ---- set by the host for the launch (mcl_engine.cu: fast_plan) -------------------------------------------------
    int frac_bits;       // fixed-point fractional bits FB (10..12: as many as the window's extent leaves room for)
    int fmask;           // (bits & fmask) == 0  <=>  the coordinate is within the uncertain band of an integer
    int mbk;             // bit pattern of the fixed-point magic number >> FB: (bits >> FB) - mbk = cell
    float magic;         // 1.5*2^(23-FB) + kb/2^FB: a float add leaves round(v*2^FB) + kb in the low mantissa bits
    float band;          // kb/2^FB + half a fixed-point step: the slack of the certain-interior box
    float t_dir;         // 3(1 + eps) + slack: octant decisions are certain beyond this
    float t_dir_neg;     // 5(1 + eps) + slack: the same when the extended point has a negative coordinate
    float x2_min;        // extended-point (global) coordinates at or above this are certainly non-negative in the reference
    float ghalf_x, ghalf_y;   // |e - gmid| >= ghalf  =>  the endpoint is certainly two or more cells outside the grid
    float rho_lo, rho_hi;   // range of the scan's interpolation ratios
    float rho_abs;          // max |ratio|
    float ang_room;         // 9.5 - max |beam angle|: |heading| + rho_abs |heading change| must stay below it
    float max_shift;     // largest |dS| (cells) a particle may have and still take the fast pass
    float coord_hi;      // largest robot cell coordinate (global) the error budget covers
    float reach;         // longest ray of the scan in cells (+ margin)
    float grid_min_dim;  // min(W, H)
    int grid_w, grid_h;
    // ---- set per window (mcl_kernels.cuh: plan_set_window).  The float model works in WINDOW-RELATIVE cell
    // coordinates: its coordinate roundings then happen at the window's magnitude instead of the map's -------------
    float shift_x, shift_y;  // window origin (global cells)
    float mid_x, half_x; // certain-interior test on the endpoint: |e - mid| < half  <=>  cell inside [lc, hc)
    float mid_y, half_y;
    float gmid_x, gmid_y;    // grid centre, window-relative
    float x2_lo_x, x2_lo_y;  // x2_min, window-relative
    float pitch_f;       // window pitch as a float (the step offset is assembled in float)
    int idx_bias;        // folds the fixed-point bias into the cell index
    int safe_idx;        // window cell (1,1): read by evaluations whose endpoint is not in the window (value discarded)
};

// Absolute error bound of fast_sincos for |a| <= 9.5 (measured on B200 over every float in the range by
// mcl_debug_fast_trig_error: 1.27e-6; tests/test_gpu_parity.py::test_fast_trig_error_bound re-measures it).
constexpr float kFastTrigErr = 2.0e-6f;

// Per-particle constants of the fast pass, in WINDOW-RELATIVE cell coordinates.
struct FastBase {
    float sxb, syb, thb;     // robot cell coordinate / heading at rho = 0 (or the pose itself when !INTERP)
    float dsx, dsy, dth;     // change over rho = 0..1
    bool ok;                 // false: every beam of this particle goes to the exact pass
    int edge;                // 0: every endpoint and doubled endpoint of this particle is inside the grid (no edge tests
                             // needed); 1: endpoints may leave the grid; 2: doubled endpoints may also turn negative
};

template <bool INTERP>
__device__ __forceinline__ FastBase make_fast_base(float xa, float ya, float tha, float xb, float yb, float thb,
                                                   double gx, double gy, double cpm_d, const FastPlan& fp)
{
    FastBase f;
    double gsx, gsy;             // robot cell coordinate in GLOBAL cells (double): validity checks, then shifted
    if (INTERP) {
        gsx = __dmul_rn(__dsub_rn((double)xb, gx), cpm_d);
        gsy = __dmul_rn(__dsub_rn((double)yb, gy), cpm_d);
        f.thb = thb;
        f.dsx = (float)__dmul_rn((double)__fsub_rn(xa, xb), cpm_d);     // the reference's float difference (interpolation.hpp:39)
        f.dsy = (float)__dmul_rn((double)__fsub_rn(ya, yb), cpm_d);
        f.dth = (float)fold_pi(__dsub_rn((double)tha, (double)thb));    // angle_diff (:41)
    } else {
        gsx = __dmul_rn(__dsub_rn((double)xa, gx), cpm_d);
        gsy = __dmul_rn(__dsub_rn((double)ya, gy), cpm_d);
        f.thb = tha;
        f.dsx = 0.0f; f.dsy = 0.0f; f.dth = 0.0f;
    }
    f.sxb = (float)__dsub_rn(gsx, (double)fp.shift_x);
    f.syb = (float)__dsub_rn(gsy, (double)fp.shift_y);
    // The robot's GLOBAL cell coordinate must stay in [1, coord_hi] over the whole sweep (truncation == floor, the
    // budget's magnitudes), the headings must be wrapped ones, and the particle must not jump: everything else (NaN
    // included: the comparisons fail) is left to the exact pass.
    const float gxb = (float)gsx, gyb = (float)gsy;
    const float x0 = __fmaf_rn(f.dsx, fp.rho_lo, gxb), x1 = __fmaf_rn(f.dsx, fp.rho_hi, gxb);
    const float y0 = __fmaf_rn(f.dsy, fp.rho_lo, gyb), y1 = __fmaf_rn(f.dsy, fp.rho_hi, gyb);
    const float lo = fminf(fminf(x0, x1), fminf(y0, y1)), hi = fmaxf(fmaxf(x0, x1), fmaxf(y0, y1));
    f.ok = lo >= 1.0f && hi <= fp.coord_hi && fabsf(f.dsx) <= fp.max_shift && fabsf(f.dsy) <= fp.max_shift &&
           fabsf(f.thb) <= 3.15f && fabsf(f.dth) <= 3.15f &&
           __fmaf_rn(fp.rho_abs, fabsf(f.dth), fabsf(f.thb)) <= fp.ang_room;      // SFU error bound: |angle| <= 9.5
    // how close to the grid's edges the particle's rays can get (reach already carries the margins)
    const bool inside = lo >= fp.reach && hi <= fp.grid_min_dim - fp.reach;      // no endpoint within 2 cells of an edge
    const bool x2_pos = lo >= 2.0f * fp.reach;                                   // no doubled endpoint below 3 cells
    f.edge = x2_pos ? (inside ? 0 : 1) : 2;
    return f;
}

// SFU sine and cosine of an angle |a| <= 9.5 rad (MUFU.SIN/COS after scaling to revolutions).
__device__ __forceinline__ void fast_sincos(float a, float* s, float* c)
{
    *s = __sinf(a);
    *c = __cosf(a);
}

#ifdef MCL_FAST_DIAG
__device__ unsigned long long g_fast_diag[8];
#endif
// The float model of one evaluation: p = range*cpm*(cos, sin), e = robot cell coordinate + p.
template <bool INTERP>
__device__ __forceinline__ void fast_endpoint(const FastBase& p, const FastBeam& b, float& px, float& py, float& ex,
                                              float& ey)
{
    const float sx = INTERP ? __fmaf_rn(p.dsx, b.ratio, p.sxb) : p.sxb;
    const float sy = INTERP ? __fmaf_rn(p.dsy, b.ratio, p.syb) : p.syb;
    const float thr = INTERP ? __fmaf_rn(p.dth, b.ratio, p.thb) : p.thb;
    float s, c;
    fast_sincos(__fsub_rn(thr, b.theta), &s, &c);
    px = __fmul_rn(b.rc, c); py = __fmul_rn(b.rc, s);
    ex = __fadd_rn(px, sx); ey = __fadd_rn(py, sy);
}

// Window cell read for the fast pass: shared-memory reads go through ld.shared.s8 on a 32-bit shared address (one LDS
// with the sign extension built in; keeps the compiler from re-deriving the value through packed 16-bit selects).
template <bool SMEM>
__device__ __forceinline__ int fast_read(const int8_t* __restrict__ cells, unsigned sbase, int idx)
{
    if (SMEM) {
        int v;
        asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(sbase + (unsigned)idx));
        return v;
    }
    return (int)__ldg(cells + idx);
}

// One certified evaluation.  Returns true when certain; v2 is then the ray's score in half units (else 0).
// Straight-line on purpose (bitwise &, no short-circuit): every lane runs the same ~60 instructions.
// EDGE (warp-uniform, from FastBase::edge): 2 = all tests; 1 = the doubled endpoint cannot be negative (its test is
// skipped); 0 = additionally no endpoint can leave the grid (the out-of-grid test is skipped).
template <bool INTERP, bool SMEM, bool COUNT, int EDGE>
__device__ __forceinline__ bool score_beam_fast(const FastBase& p, const FastBeam& b, const FastPlan& fp,
                                                const int8_t* __restrict__ cells, unsigned sbase, int pitch, int& v2,
                                                int& gathers)
{
    float px, py, ex, ey;
    fast_endpoint<INTERP>(p, b, px, py, ex, ey);
    // fixed point: low FB bits = fraction (+ the band offset), the rest = cell (+ bias)
    const int bx = __float_as_int(__fadd_rn(ex, fp.magic)), by = __float_as_int(__fadd_rn(ey, fp.magic));
    // endpoint cell certain: not within eps of a cell boundary, interior to the window and inside the grid
    const bool frac_ok = min((unsigned)(bx & fp.fmask), (unsigned)(by & fp.fmask)) != 0u;
    const bool in_x = fabsf(__fsub_rn(ex, fp.mid_x)) < fp.half_x;
    const bool in_y = fabsf(__fsub_rn(ey, fp.mid_y)) < fp.half_y;
    const bool in_win = in_x & in_y;
    const bool cell_ok = frac_ok & in_win;
    // octant class of the direction; the extended point must not have a negative coordinate
    const float ax = fabsf(px), ay = fabsf(py);
    const float d1 = __fsub_rn(__fadd_rn(ax, ax), ay);      // step x iff 2|ddx| >= |ddy|
    const float d2 = __fsub_rn(__fadd_rn(ay, ay), ax);      // step y iff |ddx| <= 2|ddy|
    // an extended point at a negative coordinate is truncated toward zero by the reference (not floored), which
    // widens the band of its differences from 3 to 5 cells
    const bool x2_pos = EDGE < 2 || ((__fadd_rn(ex, px) >= fp.x2_lo_x) & (__fadd_rn(ey, py) >= fp.x2_lo_y));
    const float t_dir = x2_pos ? fp.t_dir : fp.t_dir_neg;
    const bool dir_ok = fminf(fabsf(d1), fabsf(d2)) > t_dir;
    // endpoint certainly two or more cells outside the grid: it and both neighbours read 0 (occupancy_grid.cpp:65-70)
    const bool outside = EDGE >= 1 && ((fabsf(__fsub_rn(ex, fp.gmid_x)) >= fp.ghalf_x) |
                                       (fabsf(__fsub_rn(ey, fp.gmid_y)) >= fp.ghalf_y));
    // step toward the extended point, assembled in float: (+-1 or 0) + (+-1 or 0) * pitch, then to an integer by a
    // magic-number add (no conversion unit)
    const float ux = __uint_as_float((__float_as_uint(px) & 0x80000000u) | 0x3f800000u);
    const float uy = __uint_as_float((__float_as_uint(py) & 0x80000000u) | 0x3f800000u);
    const float offf = __fmaf_rn(d2 > 0.0f ? uy : 0.0f, fp.pitch_f, d1 > 0.0f ? ux : 0.0f);
    const int off = __float_as_int(__fadd_rn(offf, 12582912.0f)) - 0x4b400000;
    const int cidx = (int)((unsigned)(by >> fp.frac_bits) * (unsigned)pitch + (unsigned)(bx >> fp.frac_bits) -
                           (unsigned)fp.idx_bias);
    const int idx = in_win ? cidx : fp.safe_idx;       // an endpoint near a cell boundary still reads its float-pass cell
    const int odds = fast_read<SMEM>(cells, sbase, idx);
    const int o1 = fast_read<SMEM>(cells, sbase, idx - off);      // toward the robot
    const int o2 = fast_read<SMEM>(cells, sbase, idx + off);      // toward the extended point
    // odds == -1 (derived map, mcl_kernels.cuh: derive_fast_map_kernel): nothing positive within two cells of the
    // float-pass cell, hence within one cell of the reference's endpoint cell: the score is 0 whichever cell and octant
    const bool val_ok = cell_ok & ((odds > 0) | dir_ok);       // the cell reads below are the reference's
    const bool certain = val_ok | outside | (in_win & (odds == -1));
#ifdef MCL_FAST_DIAG
    {   // why evaluations are deferred (diagnostic build only)
        const bool dirband = !(fminf(fabsf(d1), fabsf(d2)) > fp.t_dir);
        const bool x2neg = !dir_ok;
        atomicAdd(&g_fast_diag[0], 1ull);
        if (!frac_ok) atomicAdd(&g_fast_diag[1], 1ull);
        if (!(in_x & in_y) && !outside) atomicAdd(&g_fast_diag[2], 1ull);
        if (cell_ok && odds <= 0 && dirband) atomicAdd(&g_fast_diag[3], 1ull);
        if (cell_ok && odds <= 0 && !dirband && x2neg) atomicAdd(&g_fast_diag[4], 1ull);
        if (!certain) atomicAdd(&g_fast_diag[5], 1ull);
    }
#endif
    if (COUNT) gathers += certain ? (val_ok & (odds > 0) ? 1 : 3) : 0;
    const int v = odds > 0 ? 2 * odds : (o1 > 0 ? o1 : max(o2, 0));
    v2 = val_ok ? v : 0;       // outside / empty neighbourhood: score 0 (an empty neighbourhood also gives v == 0)
    return certain;
}

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter-based: counter = (global particle index, update number), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// Four uniform words -> up to four standard normals (Box-Muller, float).
__device__ __forceinline__ void philox_normals(uint4 w, float& z0, float& z1, float& z2, float& z3)
{
    const float k = 2.3283064365386963e-10f;   // 2^-32
    const float u0 = fmaf((float)w.x, k, 0.5f * k), u1 = (float)w.y * k;
    const float u2 = fmaf((float)w.z, k, 0.5f * k), u3 = (float)w.w * k;
    const float r0 = sqrtf(-2.0f * logf(fminf(u0, 0.99999994f)));
    const float r1 = sqrtf(-2.0f * logf(fminf(u2, 0.99999994f)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    z0 = r0 * c0; z1 = r0 * s0; z2 = r1 * c1; z3 = r1 * s1;
}

}  // namespace mcl
