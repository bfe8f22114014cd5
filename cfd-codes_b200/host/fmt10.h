/*
 * host/fmt10.h -- exact, fast "%.10f" for the Tecplot writer (the reference's datwrite prints
 * 7 * N^3 values with fprintf("%.10f"), main.c:376 -- 117 M conversions at 256^3).
 *
 * fmt10(buf, x) writes exactly the characters printf("%.10f", x) produces (glibc: the exact binary
 * value rounded half-to-even at the 10th decimal, sign kept for -0.0 and for negative values that
 * round to zero) and returns their count.  Integer arithmetic only: |x| = m * 2^e with a 53-bit m,
 * so |x| * 10^10 = (m * 10^10) >> -e is computed exactly in 128 bits.  Values outside
 * [0, 1e8) and non-finite values fall back to snprintf.
 */
#ifndef LBM_FMT10_H
#define LBM_FMT10_H
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline int fmt10(char *buf, double x) {
    uint64_t bits;
    memcpy(&bits, &x, sizeof bits);
    const int neg = (int)(bits >> 63);
    const int bexp = (int)((bits >> 52) & 0x7ff);
    uint64_t m = bits & 0xfffffffffffffULL;
    if (bexp == 0x7ff || fabs(x) >= 1e8) return sprintf(buf, "%.10f", x);
    int e;
    if (bexp == 0) {
        e = -1074; /* subnormal (or zero) */
    } else {
        m |= 1ULL << 52;
        e = bexp - 1075;
    }
    unsigned __int128 num = (unsigned __int128)m * 10000000000ULL; /* < 2^87 */
    uint64_t r;
    if (e >= 0) {
        r = (uint64_t)(num << e); /* |x| < 1e8: fits */
    } else {
        const int s = -e;
        if (s >= 127) {
            r = 0; /* |x| * 1e10 < 2^-40 */
        } else {
            const unsigned __int128 q = num >> s;
            const unsigned __int128 rem = num - (q << s);
            const unsigned __int128 half = (unsigned __int128)1 << (s - 1);
            r = (uint64_t)q;
            if (rem > half || (rem == half && (r & 1))) r++;
        }
    }
    const uint64_t ip = r / 10000000000ULL;
    uint64_t fp = r % 10000000000ULL;
    char *p = buf;
    if (neg) *p++ = '-';
    char tmp[24];
    int n = 0;
    uint64_t t = ip;
    do { tmp[n++] = (char)('0' + t % 10); t /= 10; } while (t);
    while (n) *p++ = tmp[--n];
    *p++ = '.';
    for (int d = 9; d >= 0; d--) { p[d] = (char)('0' + fp % 10); fp /= 10; }
    p += 10;
    *p = 0;
    return (int)(p - buf);
}
#endif
