/* Test helper: compares botlab_b200/csrc/glibc_sincosf.h with the live libm sincosf, bit for bit.
 * usage: sincosf_sweep <max_abs_as_float> <stride>   -> prints "checked N mismatches M" */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../../botlab_b200/csrc/glibc_sincosf.h"

int main(int argc, char** argv)
{
    float maxabs = argc > 1 ? (float)atof(argv[1]) : 4.0f;
    uint32_t stride = argc > 2 ? (uint32_t)atoi(argv[2]) : 1;
    uint32_t top;
    memcpy(&top, &maxabs, 4);
    unsigned long long checked = 0, bad = 0;
    #pragma omp parallel for reduction(+:checked,bad) schedule(static)
    for (uint32_t b = 0; b <= top; b += stride) {
        for (int sgn = 0; sgn < 2; ++sgn) {
            uint32_t u = b | ((uint32_t)sgn << 31);
            float x, s0, c0, s1, c1;
            memcpy(&x, &u, 4);
            sincosf(x, &s0, &c0);
            glibc_sincosf(x, &s1, &c1);
            float s2, c2;
            glibc_sincosf_fast(x, &s2, &c2);
            ++checked;
            if (memcmp(&s0, &s1, 4) || memcmp(&c0, &c1, 4) || memcmp(&s0, &s2, 4) || memcmp(&c0, &c2, 4)) {
                if (bad < 5) fprintf(stderr, "mismatch x=%a sin %a vs %a cos %a vs %a\n", x, s0, s1, c0, c1);
                ++bad;
            }
        }
    }
    /* wrap fast path: every float a in (-3*pi, -pi] */
    if (stride == 1 || argc > 3) {
        float lo = -9.42477f, hi = -3.14159274f;
        uint32_t ulo, uhi;
        memcpy(&ulo, &lo, 4);
        memcpy(&uhi, &hi, 4);
        #pragma omp parallel for reduction(+:checked,bad) schedule(static)
        for (uint32_t u = uhi; u <= ulo; ++u) {        /* negative floats: larger bits = more negative */
            float a;
            memcpy(&a, &u, 4);
            float want = (float)((double)a + 2.0 * M_PI);
            float t = a + GS_TWO_PI_HI;
            float got = (fabsf(t) >= 9.5367431640625e-07f) ? t + GS_TWO_PI_LO : want;
            ++checked;
            if (memcmp(&want, &got, 4)) { if (bad < 5) fprintf(stderr, "wrap mismatch a=%a\n", a); ++bad; }
        }
    }
    printf("checked %llu mismatches %llu\n", checked, bad);
    return bad != 0;
}
