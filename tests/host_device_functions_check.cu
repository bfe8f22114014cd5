// tests/host_device_functions_check.cu -- TEST HARNESS (CPU): compiles the PRODUCT's device functions
// lbm3d::collide<METHOD, G1> and lbm3d::apply_boundary (csrc/lbm3d_kernels.cuh, d3q19_rules.h) for the
// host and checks them bit-for-bit against the oracle's single-cell entry points on random cells:
// every boundary class (26) and both collision operators, with and without preconditioning.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -fmad=false -Xcompiler -ffp-contract=off \
//        -I cfd-codes_b200/csrc tests/host_device_functions_check.cu -Loracle -llbm_oracle -o check && ./check
// (test infrastructure: the library exports no host compute path; this file is never shipped)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "lbm3d_kernels.cuh"
#include "lbm2d_kernels.cuh"

extern "C" {
struct oracle3d_params { int N, Re, METHOD, DIAG; double UMAX, GAMMA, MAGIC; };
void *oracle3d_create(const oracle3d_params *);
void oracle3d_destroy(void *);
void oracle3d_derived(void *, double *out8);
int oracle3d_collide_cell(void *, const double *f19, double *out19);
int oracle3d_rule_cell(void *, int bx, int by, int bz, double *f19);
struct oracle2d_params { int N, M; double Re, Ulid, MAGIC; };
void oracle2d_collide_cell(const oracle2d_params *, const double *ft9, double *out9);
void oracle2d_derived(const oracle2d_params *, double *out4);
}

static unsigned long long rng = 0x9E3779B97F4A7C15ULL;
static double rnd() {
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    return (double)(rng >> 11) * (1.0 / 9007199254740992.0);
}

template <int METHOD, bool G1>
static int check_collide(void *h, const lbm3d::KP &p, int n) {
    int bad = 0;
    for (int t = 0; t < n; t++) {
        double f[19], g[19], want[19];
        const double w[19] = {1.0 / 3, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 18, 1.0 / 36, 1.0 / 36, 1.0 / 36,
                              1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36};
        const double amp = (t % 7 == 0) ? 0.9 : 0.08;   // occasionally far from equilibrium (negative feq / populations)
        for (int q = 0; q < 19; q++) f[q] = w[q] * (1.0 + amp * (2.0 * rnd() - 1.0));
        memcpy(g, f, sizeof g);
        const int neg_o = oracle3d_collide_cell(h, f, want);
        const bool neg_p = lbm3d::collide<METHOD, G1>(g, p);
        if (memcmp(g, want, sizeof g) != 0 || (neg_o != 0) != neg_p) {
            if (bad < 3) printf("collide<%d,%d> mismatch at sample %d (neg %d vs %d)\n", METHOD, (int)G1, t, neg_o, (int)neg_p);
            bad++;
        }
    }
    return bad;
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 20000;
    int bad = 0;
    const double gammas[3] = {1.0, 0.5, 2.0};
    for (int method = 0; method < 2; method++)
        for (int diag = 0; diag < 2; diag++)
            for (int gi = 0; gi < 3; gi++) {
                oracle3d_params op = {24, 100, method, diag, 0.1, gammas[gi], 0.25};
                void *h = oracle3d_create(&op);
                double d[8];
                oracle3d_derived(h, d);   // NU TAU OMEGA OMEGAm OmOMEGA ulid vlid wlid
                lbm3d::KP p;
                memset(&p, 0, sizeof p);
                p.omega = d[2]; p.omegam = d[3]; p.omomega = d[4]; p.gammainv = 1.0 / gammas[gi];
                p.ulid = d[5]; p.vlid = d[6]; p.wlid = d[7];
                p.vfrac = 1.0 / (p.vlid + 1.0); p.third = 1.0 / 3.0; p.sixth = 1.0 / 6.0;
                if (method == 1) {
                    bad += check_collide<1, false>(h, p, n);
                    if (gi == 0) bad += check_collide<1, true>(h, p, n);
                } else {
                    bad += check_collide<0, false>(h, p, n);
                    if (gi == 0) bad += check_collide<0, true>(h, p, n);
                }
                // every boundary class
                for (int bx = 0; bx < 3; bx++)
                    for (int by = 0; by < 3; by++)
                        for (int bz = 0; bz < 3; bz++) {
                            if (bx + by + bz == 0) continue;
                            for (int t = 0; t < n / 20; t++) {
                                double f[19], g[19];
                                for (int q = 0; q < 19; q++) f[q] = 0.01 + 0.06 * rnd();
                                memcpy(g, f, sizeof g);
                                const int kind = oracle3d_rule_cell(h, bx, by, bz, f);
                                lbm3d::apply_boundary(g, LBM_CLS(bx, by, bz), p);
                                if (kind <= 0 || memcmp(f, g, sizeof f) != 0) {
                                    if (bad < 6) printf("rule mismatch class (%d,%d,%d) kind %d\n", bx, by, bz, kind);
                                    bad++;
                                }
                            }
                        }
                oracle3d_destroy(h);
            }
    // D2Q9 collision (lbm2d::collide) against the 2D oracle's, three relaxation times
    for (int cfg = 0; cfg < 3; cfg++) {
        oracle2d_params op = {100, 300, cfg == 0 ? 100.0 : (cfg == 1 ? 1000.0 : 25.0), 0.1, cfg == 2 ? 0.1875 : 0.25};
        double d[4];
        oracle2d_derived(&op, d);   // NU TAU OMEGA OMEGAm
        lbm2d::KP2 p;
        memset(&p, 0, sizeof p);
        p.omega = d[2]; p.homega = 0.5 * d[2]; p.homegam = 0.5 * d[3];
        for (int t = 0; t < n; t++) {
            double ft[9], want[9];
            for (int k = 0; k < 9; k++) ft[k] = 0.02 * (2.0 * rnd() - 1.0);   // deviation form: f - w
            oracle2d_collide_cell(&op, ft, want);
            lbm2d::collide(ft, p);
            if (memcmp(ft, want, sizeof ft) != 0) { if (bad < 9) printf("D2Q9 collide mismatch (cfg %d)\n", cfg); bad++; }
        }
    }
    printf(bad ? "FAILED (%d)\n" : "host check of the product's device functions ok\n", bad);
    return bad ? 1 : 0;
}
