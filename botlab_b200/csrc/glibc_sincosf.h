// Bit-exact restatement of glibc 2.39's sincosf for |x| < 120, usable from host C/C++ and from CUDA device code.
//
// Why: the reference's sensor model truncates `range * cosf(theta) * cellsPerMeter + start` to a cell index
// (src/slam/sensor_model.cpp:34-38), so a 1-ulp difference in cosf/sinf can move an endpoint to another cell and change
// a particle's score by whole units.  g++ -O3 merges the reference's std::cos/std::sin(float) into one `sincosf` call,
// and CUDA's sincosf is a different (also not correctly rounded) function.  Parity therefore needs glibc's algorithm:
// third-party code, not in /root/reference: glibc 2.39 (Ubuntu 2.39-0ubuntu8.5), sysdeps/ieee754/flt-32/s_sincosf.c
// with the x86-64 FMA ifunc variant (__sincosf_fma, sysdeps/x86_64/fpu/multiarch + sincosf_poly.h) -- the one every
// FMA-capable x86-64 host selects.  The algorithm (Szabolcs Nagy's ARM optimized-routines sincosf): reduce
// x to [-pi/4, pi/4] in double with n = round(x * 2/pi), evaluate degree-7/8 double polynomials, round once to float.
// The operation order and fusion below were read off the disassembly of libm.so.6's __sincosf_fma and the constants
// from its .rodata; tests/test_sincosf.py checks it against the live libm over every float in [-4, 4] (strided in CI).
#ifndef BOTLAB_B200_GLIBC_SINCOSF_H
#define BOTLAB_B200_GLIBC_SINCOSF_H

#include <stdint.h>
#if defined(__CUDA_ARCH__)
#define GS_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GS_MUL(a, b) __dmul_rn((a), (b))
#define GS_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define GS_FMA(a, b, c) fma((a), (b), (c))
#define GS_MUL(a, b) ((a) * (b))
#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD static inline
#endif
#endif

// Polynomial coefficients (__sincosf_table[0]; table[1] negates the cosine set).
#define GS_C0 1.0
#define GS_C1 (-0x1.ffffffd0c621cp-2)
#define GS_C2 0x1.55553e1068f19p-5
#define GS_C3 (-0x1.6c087e89a359dp-10)
#define GS_C4 0x1.99343027bf8c3p-16
#define GS_S1 (-0x1.555545995a603p-3)
#define GS_S2 0x1.1107605230bc4p-7
#define GS_S3 (-0x1.994eb3774cf24p-13)
#define GS_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define GS_HPI 0x1.921FB54442D18p0       /* pi/2 */

GS_HD uint32_t gs_float_bits(float x)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
#endif
}

// Valid for |x| < 120 (the engine only ever passes angles wrapped to [-pi, pi]); larger |x| would need glibc's
// reduce_large table, which this path never reaches.
GS_HD void glibc_sincosf(float xf, float* sinp, float* cosp)
{
    const uint32_t top = (gs_float_bits(xf) >> 20) & 0x7ff;
    double x = (double)xf;
    double xs, x2, csign;   // xs: reduced argument with the quadrant sign folded in; csign: +1/-1 on the cosine set
    int swap;
    if (top < 0x3f4) {                      // |x| < pi/4 (abstop12 compare)
        if (top < 0x398) {                  // |x| < 2^-12
            *sinp = xf;
            *cosp = 1.0f;
            return;
        }
        xs = x;
        x2 = GS_MUL(x, x);
        csign = 1.0;
        swap = 0;
    } else {
        double r = GS_MUL(x, GS_HPI_INV);
        int n = ((int32_t)r + 0x800000) >> 24;              // round to nearest quadrant
        double xr = GS_FMA(-(double)n, GS_HPI, x);          // vfnmadd: x - n*hpi, fused
        double sgn = ((n + 1) & 2) ? -1.0 : 1.0;            // sign[n & 3] = {1,-1,-1,1}
        xs = GS_MUL(xr, sgn);
        x2 = GS_MUL(xr, xr);
        csign = (n & 2) ? -1.0 : 1.0;
        swap = n & 1;
    }
    double x3 = GS_MUL(x2, xs);
    double x4 = GS_MUL(x2, x2);
    double s1 = GS_FMA(x2, GS_S3, GS_S2);
    double c2 = GS_FMA(x2, csign * GS_C4, csign * GS_C3);
    double c1 = GS_FMA(x2, csign * GS_C1, csign * GS_C0);
    double x5 = GS_MUL(x2, x3);
    double x6 = GS_MUL(x2, x4);
    double s = GS_FMA(x3, GS_S1, xs);
    double c = GS_FMA(x4, csign * GS_C2, c1);
    float sv = (float)GS_FMA(x5, s1, s);
    float cv = (float)GS_FMA(x6, c2, c);
    if (swap) { *sinp = cv; *cosp = sv; }
    else      { *sinp = sv; *cosp = cv; }
}

// Branch-free variant for the hot loop, bit-identical to glibc_sincosf for |x| < 120 (checked exhaustively on [-4, 4]
// by tests/csrc/sincosf_sweep.c):
//  - the quadrant formula n = ((int)(x*hpi_inv) + 2^23) >> 24 gives n = 0 for |x| < pi/4, where x - 0*hpi = x exactly,
//    so glibc's separate small-argument path computes the same polynomial;
//  - for |x| < 2^-12 the polynomial itself rounds to (x, 1.0f), glibc's shortcut values;
//  - __sincosf_table[1] is table[0] with the cosine coefficients negated, and sign[n&3] multiplies the reduced
//    argument of an odd polynomial: round-to-nearest is sign-symmetric, so both signs can be applied to the final
//    floats instead (no multiplies by +-1).
#if defined(__CUDACC__)
// Device copies of the constants in constant memory: FP64 instructions take them as c[bank][offset] operands, instead of
// the two 32-bit immediate moves per use that literal doubles cost.
__device__ __constant__ double gs_k[10] = {GS_HPI_INV, GS_HPI, GS_S1, GS_S2, GS_S3, GS_C0, GS_C1, GS_C2, GS_C3, GS_C4};
#endif
#if defined(__CUDA_ARCH__)
#define GS_K(i, lit) gs_k[i]
#else
#define GS_K(i, lit) (lit)
#endif

#if defined(__CUDACC__)
// The hot loop keeps the ten constants in registers: gs_load_consts() reads them once per thread through volatile
// loads, which the compiler cannot rematerialise (as immediates or constant-bank loads) inside the loop.
struct GsConsts { double hpi_inv, hpi, s1, s2, s3, c0, c1, c2, c3, c4; };
__device__ double gs_kg[10] = {GS_HPI_INV, GS_HPI, GS_S1, GS_S2, GS_S3, GS_C0, GS_C1, GS_C2, GS_C3, GS_C4};
__device__ __forceinline__ GsConsts gs_load_consts()
{
#if !defined(MCL_PIN_CONSTS) || MCL_PIN_CONSTS
    // volatile global loads: neither NVVM nor ptxas may re-issue them, so the values stay in registers
    const volatile double* p = gs_kg;
    GsConsts k = {p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9]};
#else
    // constant-bank operands: ten fewer 64-bit registers, for builds that trade them for resident warps
    GsConsts k = {gs_k[0], gs_k[1], gs_k[2], gs_k[3], gs_k[4], gs_k[5], gs_k[6], gs_k[7], gs_k[8], gs_k[9]};
#endif
    return k;
}
// Same arithmetic as glibc_sincosf_core with the constants passed in.
__device__ __forceinline__ void glibc_sincosf_regs(const GsConsts& k, float xf, float* sinp, float* cosp)
{
    const double x = (double)xf;
    const double r = __dmul_rn(x, k.hpi_inv);
    const int n = ((int32_t)r + 0x800000) >> 24;
    const double xr = __fma_rn(-(double)n, k.hpi, x);
    const double x2 = __dmul_rn(xr, xr);
    const double x3 = __dmul_rn(x2, xr);
    const double x4 = __dmul_rn(x2, x2);
    const double s1 = __fma_rn(x2, k.s3, k.s2);
    const double c2 = __fma_rn(x2, k.c4, k.c3);
    const double c1 = __fma_rn(x2, k.c1, k.c0);
    const double x5 = __dmul_rn(x2, x3);
    const double x6 = __dmul_rn(x2, x4);
    const double s = __fma_rn(x3, k.s1, xr);
    const double c = __fma_rn(x4, k.c2, c1);
    const float sv = (float)__fma_rn(x5, s1, s);
    const float cv = (float)__fma_rn(x6, c2, c);
    const uint32_t sbit = ((uint32_t)(n + 1) & 2u) << 30;
    const uint32_t cbit = ((uint32_t)n & 2u) << 30;
    const float ss = __uint_as_float(__float_as_uint(sv) ^ sbit);
    const float cc = __uint_as_float(__float_as_uint(cv) ^ cbit);
    *sinp = (n & 1) ? cc : ss;
    *cosp = (n & 1) ? ss : cc;
}
#endif

// _core differs from glibc in exactly one input: x = -0.0f yields sin = +0.0f instead of -0.0f (the polynomial's
// x3*S1 term is +0).  The sensor model only ever truncates range*sin*cpm + start to an int, where the sign of a zero
// cannot matter, so the hot loop uses _core; glibc_sincosf_fast adds the one select that makes it exact everywhere.
GS_HD void glibc_sincosf_core(float xf, float* sinp, float* cosp)
{
    const double x = (double)xf;
    const double r = GS_MUL(x, GS_K(0, GS_HPI_INV));
    const int n = ((int32_t)r + 0x800000) >> 24;
    const double xr = GS_FMA(-(double)n, GS_K(1, GS_HPI), x);
    const double x2 = GS_MUL(xr, xr);
    const double x3 = GS_MUL(x2, xr);
    const double x4 = GS_MUL(x2, x2);
    const double s1 = GS_FMA(x2, GS_K(4, GS_S3), GS_K(3, GS_S2));
    const double c2 = GS_FMA(x2, GS_K(9, GS_C4), GS_K(8, GS_C3));
    const double c1 = GS_FMA(x2, GS_K(6, GS_C1), GS_K(5, GS_C0));
    const double x5 = GS_MUL(x2, x3);
    const double x6 = GS_MUL(x2, x4);
    const double s = GS_FMA(x3, GS_K(2, GS_S1), xr);
    const double c = GS_FMA(x4, GS_K(7, GS_C2), c1);
    const float sv = (float)GS_FMA(x5, s1, s);
    const float cv = (float)GS_FMA(x6, c2, c);
    const uint32_t sbit = ((uint32_t)(n + 1) & 2u) << 30;    // sign[n & 3] = {+,-,-,+}
    const uint32_t cbit = ((uint32_t)n & 2u) << 30;          // table[1] when n & 2
#if defined(__CUDA_ARCH__)
    const float ss = __uint_as_float(__float_as_uint(sv) ^ sbit);
    const float cc = __uint_as_float(__float_as_uint(cv) ^ cbit);
#else
    uint32_t us = gs_float_bits(sv) ^ sbit, uc = gs_float_bits(cv) ^ cbit;
    float ss, cc;
    memcpy(&ss, &us, 4);
    memcpy(&cc, &uc, 4);
#endif
    *sinp = (n & 1) ? cc : ss;
    *cosp = (n & 1) ? ss : cc;
}

GS_HD void glibc_sincosf_fast(float xf, float* sinp, float* cosp)
{
    float s, c;
    glibc_sincosf_core(xf, &s, &c);
    *sinp = (xf == 0.0f) ? xf : s;
    *cosp = c;
}

// wrap_to_pi (common/angle_functions.hpp:12-24) fast path for a in (-3*pi, -pi]: the reference computes
// (float)((double)a + 2*M_PI).  With 2*M_PI = H + L (H = (float)(2*M_PI)), a + H is exact in float (Sterbenz range) and
// the double sum is exact too, so the result is RN_float(t + L) with t = a + H.  Adding the float-rounded L instead gives
// the same rounding unless |t| is tiny (checked exhaustively by tests/csrc/sincosf_sweep.c over every float in range);
// callers take the double path when |t| < 2^-20.
#define GS_TWO_PI_HI 6.2831854820251465f       /* (float)(2*M_PI) */
#define GS_TWO_PI_LO (-1.7484555e-7f)          /* (float)(2*M_PI - H) */

#endif  // BOTLAB_B200_GLIBC_SINCOSF_H
