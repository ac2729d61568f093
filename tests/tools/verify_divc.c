/* verify_divc.c -- test tool: is  q=x*rc; r=fma(-c,q,x); q'=fma(r,rc,q)  (div_const in
 * env_build_b200/csrc/ce2e_device.cuh) bit-identical to the IEEE quotient x/c ?
 * Usage: verify_divc [stride]   stride 1 = every fp32 bit pattern (about a minute per divisor).
 * Domain checked: finite x whose quotient is a normal number and where x*rc does not overflow.
 * Prints "mismatches=N" per divisor (N must be 0) and the count of out-of-domain differences. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char **argv) {
    uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    const float PI32 = 3.14159274101257324f;
    /* -15.625: the right-turn arc enters the kernel as a signed radius (x / -R == -(x / R)) */
    const float divisors[6] = {180.0f, PI32, 10.0f, 26.875f, 15.625f, -15.625f};
    int bad = 0;
    for (int d = 0; d < 6; ++d) {
        const float c = divisors[d];
        const volatile float one = 1.0f;
        const float rc = one / c;
        uint64_t mism = 0, outside = 0, n = 0;
        for (uint64_t b = 0; b < (1ull << 32); b += stride) {
            float x = from_bits((uint32_t)b);
            if (!isfinite(x)) continue;
            float want = x / c;
            float q = x * rc;
            float r = fmaf(-c, q, x);
            float got = fmaf(r, rc, q);
            ++n;
            if (to_bits(got) == to_bits(want)) continue;
            if (!isfinite(q) || fabsf(want) < 1.17549435e-38f * 16777216.0f) { ++outside; continue; }
            if (mism < 5) printf("  c=%g x=%a want=%a got=%a\n", c, x, want, got);
            ++mism;
        }
        printf("c=%.9g rc=%.9g checked=%llu mismatches=%llu outside_domain_diffs=%llu\n", c, rc,
               (unsigned long long)n, (unsigned long long)mism, (unsigned long long)outside);
        if (mism) bad = 1;
    }
    if (bad) printf("FAIL\n");
    return bad;
}
