/* tests/fmt10_check.c -- compares host/fmt10.h with printf("%.10f") on edge cases and random doubles. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fmt10.h"

static uint64_t rng = 88172645463325252ULL;
static uint64_t next(void) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; }

static int check(double x) {
    char a[400], b[400];
    fmt10(a, x);
    snprintf(b, sizeof b, "%.10f", x);
    if (strcmp(a, b)) { printf("MISMATCH %a: fmt10=%s printf=%s\n", x, a, b); return 1; }
    return 0;
}

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    int bad = 0;
    const double edge[] = {0.0, -0.0, 1.0, -1.0, 0.5, 0.00000000005, 0.00000000015, 0.00000000025, -0.00000000005, 1e-11, -1e-11,
                           4.9e-324, -4.9e-324, 2.2250738585072014e-308, 0.1, 0.07071067811865475, 255.0, 1023.0, 99999999.0,
                           123456.78901234567, 0.99999999995, 0.999999999949999, 9.99999999995, 1e8, -1e9, 1e300};
    for (size_t i = 0; i < sizeof edge / sizeof *edge; i++) bad += check(edge[i]);
    for (long k = 0; k < 2000; k++) { /* exact ties at the 10th decimal that are representable: k / 2^11 etc. */
        bad += check((double)k / 2048.0);
        bad += check(((double)k + 0.5) / 1024.0 * 1e-3);
    }
    for (long i = 0; i < n && bad < 10; i++) {
        uint64_t u = next();
        double x;
        switch (i & 3) {
            case 0: x = (double)(u >> 11) * (1.0 / 9007199254740992.0) * 0.2 - 0.1; break;          /* velocities */
            case 1: x = (double)(u >> 11) * (1.0 / 9007199254740992.0) * 1e-9; break;               /* near the last digit */
            case 2: { uint64_t b = (u & 0x800fffffffffffffULL) | ((uint64_t)(900 + (u >> 52) % 160) << 52); memcpy(&x, &b, 8); } break;
            default: x = (double)(int64_t)(u % 2000000) / 1e10 + (double)(u >> 40) * 1e-17; break;  /* near ties */
        }
        bad += check(x);
    }
    printf(bad ? "FAILED\n" : "fmt10 ok\n");
    return bad ? 1 : 0;
}
