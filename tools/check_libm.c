/*
 * check_libm.c -- exhaustive host check that rust-modem_b200/csrc/libm_f32.h is
 * bit-identical to this machine's glibc sinf/cosf/logf.
 *
 *   gcc -O2 -mfma -fopenmp -ffp-contract=off -DMG_LIBM_CONTRACT=1 \
 *       -I rust-modem_b200/csrc tools/check_libm.c -o /tmp/check_libm -lm && /tmp/check_libm
 *
 * Sweeps every binary32 in [0, hi] (default hi = 128.0, i.e. the whole |y| < 120
 * branch plus the start of the large-argument branch), the same range negated, a
 * strided sweep of all larger finite floats, and every binary32 in (0, 1] for logf
 * (the Box-Muller domain) plus a strided sweep above 1.
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "libm_f32.h"

static float asf(uint32_t u) { return mg_asfloat_host(u); }

int main(int argc, char** argv)
{
    float hi = argc > 1 ? (float)atof(argv[1]) : 128.0f;
    uint32_t hi_bits = mg_asuint_host(hi);
    long bad_sin = 0, bad_cos = 0;
    long n_trig = 0;

#pragma omp parallel for reduction(+ : bad_sin, bad_cos, n_trig) schedule(static)
    for (uint32_t u = 0; u <= hi_bits; ++u) {
        for (int neg = 0; neg < 2; ++neg) {
            float y = asf(u | ((uint32_t)neg << 31));
            float s, c;
            mg_sincosf(y, &s, &c);
            float rs = sinf(y), rc = cosf(y);
            bad_sin += mg_asuint_host(s) != mg_asuint_host(rs);
            bad_cos += mg_asuint_host(c) != mg_asuint_host(rc);
            n_trig++;
        }
    }
    /* large arguments: every 257th float up to FLT_MAX */
#pragma omp parallel for reduction(+ : bad_sin, bad_cos, n_trig) schedule(static)
    for (uint32_t u = hi_bits; u < 0x7f800000u; u += 257) {
        float y = asf(u);
        float s, c;
        mg_sincosf(y, &s, &c);
        bad_sin += mg_asuint_host(s) != mg_asuint_host(sinf(y));
        bad_cos += mg_asuint_host(c) != mg_asuint_host(cosf(y));
        n_trig++;
    }
    printf("sincosf: %ld inputs, sin mismatches %ld, cos mismatches %ld\n", n_trig, bad_sin, bad_cos);
    {
        long bad07 = 0, n07 = 0;
        const uint32_t top = mg_asuint_host(7.0f);
#pragma omp parallel for reduction(+ : bad07, n07) schedule(static)
        for (uint32_t u = 0; u <= top; ++u) {
            float y = asf(u), s, c;
            mg_sincosf_0_7(y, &s, &c);
            bad07 += mg_asuint_host(s) != mg_asuint_host(sinf(y)) || mg_asuint_host(c) != mg_asuint_host(cosf(y));
            n07++;
        }
        printf("sincosf_0_7 (branch-free form, every binary32 in [0, 7]): %ld inputs, %ld mismatches\n", n07, bad07);
    }

    long bad120 = 0, n120 = 0;
    {
        const uint32_t top = mg_asuint_host(120.0f); /* every binary32 with |y| < 120, both signs, plus the dispatching wrapper beyond */
#pragma omp parallel for reduction(+ : bad120, n120) schedule(static)
        for (uint32_t u = 0; u <= top + 4096; ++u) {
            for (int neg = 0; neg < 2; ++neg) {
                float y = asf(u | ((uint32_t)neg << 31)), s, c;
                if (u < top) mg_sincosf_lt120(y, &s, &c);
                else mg_sincosf_nco(y, &s, &c);
                bad120 += mg_asuint_host(s) != mg_asuint_host(sinf(y)) || mg_asuint_host(c) != mg_asuint_host(cosf(y));
                n120++;
            }
        }
        printf("sincosf_lt120 (branch-free form, every binary32 with |y| < 120, both signs): %ld inputs, %ld mismatches\n", n120, bad120);
    }

    printf("MG_LIBM_CONTRACT=%d %s\n", MG_LIBM_CONTRACT, (bad_sin | bad_cos | bad120) ? "MISMATCH" : "BIT-EXACT");
    return (bad_sin | bad_cos | bad120) ? 1 : 0;
}
