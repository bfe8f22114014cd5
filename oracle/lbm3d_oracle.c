/*
 * oracle/lbm3d_oracle.c -- TEST INFRASTRUCTURE ONLY (the CPU oracle).  Not product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this.  The product path (cfd-codes_b200/csrc) never links, imports or calls it.
 *
 * A plain-C restatement of the hot path of the reference program
 *   /root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c   ("3D:" below)
 * i.e. macrocollideandstream (3D:215-321), HechtBC (3D:385-675), diff (3D:678-685),
 * macrovars (3D:695-719), calcvmag (3D:730-738), initialize (3D:172-206).
 *
 * Why a restatement exists beside oracle/_ref (the real reference, compiled):
 *   - the reference indexes with `int` (NQ overflows for N >= 484, 3D:25-28,48) and allocates
 *     four lattices (3D:137-140); this file uses size_t and two lattices;
 *   - parameters are run-time values here (the reference needs a rebuild per N/Re/...);
 *   - the volume loop runs k-outer / i-inner (memory order) instead of i-outer / k-inner
 *     (3D:264-266).  Every (cell, q) slot of f2 has exactly one writer per phase, so the loop
 *     order does not change a single bit.
 * PINNING: tests/test_oracle.py checks this file bit-for-bit against oracle/_ref binaries for
 * every configuration in oracle/build_ref.py:CONFIGS and numerically against the reference's
 * golden file `example output/Solution_n=29Re=100_DIAG=0_2RT_M=0.250.dat`.
 *
 * Floating-point rules: compile with -ffp-contract=off; no re-association.
 * Representation: exactly the reference's -- two buffers f1 (current) / f2 (next), push
 * streaming into f2 only where the target is inside the box, HechtBC applied to f2 in place,
 * pointer swap.  Slots that streaming never reaches simply keep what the buffer held two steps
 * earlier -- which reproduces the buried constants and the stale f14 read at
 * x = N-1 (3D:421) without special code.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 19

/* lattice tables, 3D:27-38 */
static const double W0 = 1.0 / 3.0, W1 = 1.0 / 18.0, W2 = 1.0 / 36.0;
static const int CX[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0};
static const int CY[Q] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1};
static const int CZ[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1};

typedef struct {
    int N, Re, METHOD, DIAG;
    double UMAX, GAMMA, MAGIC;
} oracle3d_params;

typedef struct {
    oracle3d_params p;
    double w[Q];
    double NU, TAU, OMEGA, OMEGAm, OmOMEGA, gammainv, third, sixth; /* 3D:19-23,30,40-42 */
    double ulid, vlid, wlid;                                          /* 3D:183-191 */
    size_t N1, N2, N3, NQ;
    double *f1, *f2;
    int unstable;     /* checkfeq (3D:687-692) would have exit(1)'d */
    long steps_done;
    double last_diff; /* diff(f2, f1, NQ) of the most recent step, 3D:96/112 */
} oracle3d;

static inline size_t LU4(const oracle3d *o, size_t i, size_t j, size_t k, size_t q) {
    return q * o->N3 + k * o->N2 + j * o->N1 + i; /* 3D:48 with 64-bit arithmetic */
}

/* 3D:163-169 */
static inline double calcfeq(const oracle3d *o, int c1, int c2, int c3, double w, double u1,
                             double u2, double u3, double rho, double U2, int *neg) {
    double cdotu = c1 * u1 + c2 * u2 + c3 * u3;
    double feq = w * rho * (1.0 + 3.0 * cdotu + (4.5 * (cdotu * cdotu) - 1.5 * U2) * o->gammainv);
    if (feq < 0.0) *neg = 1; /* checkfeq 3D:687-692 */
    return feq;
}

void *oracle3d_create(const oracle3d_params *p) {
    oracle3d *o = (oracle3d *)calloc(1, sizeof(oracle3d));
    o->p = *p;
    o->w[0] = W0;
    for (int q = 1; q < 7; q++) o->w[q] = W1;
    for (int q = 7; q < Q; q++) o->w[q] = W2;
    /* derived constants in macro order, 3D:19-23,30 */
    o->NU = (p->UMAX) * (double)(p->N + 1.0) / (double)(p->Re);
    o->TAU = (o->NU) * (3.0 / (double)p->GAMMA) + 0.5;
    o->OMEGA = 1.0 / (double)(o->TAU);
    o->OMEGAm = 1.0 / (0.5 + (p->MAGIC / ((1.0 / o->OMEGA) - 0.5)));
    o->OmOMEGA = 1.0 - o->OMEGA;
    o->gammainv = 1.0 / p->GAMMA;
    o->sixth = 1.0 / 6.0;
    o->third = 1.0 / 3.0;
    if (p->DIAG == 1) { /* 3D:183-191 */
        o->ulid = (double)p->UMAX / sqrt(2);
        o->wlid = o->ulid;
    } else {
        o->ulid = p->UMAX;
        o->wlid = 0.0;
    }
    o->vlid = 0.0;
    o->N1 = (size_t)p->N;
    o->N2 = o->N1 * o->N1;
    o->N3 = o->N2 * o->N1;
    o->NQ = o->N3 * Q;
    o->f1 = (double *)malloc(o->NQ * sizeof(double));
    o->f2 = (double *)malloc(o->NQ * sizeof(double));
    if (!o->f1 || !o->f2) {
        free(o->f1); free(o->f2); free(o);
        return NULL;
    }
    /* initialize(), 3D:192-205: f = feq(rho = RHO0 = 1, u = 0) in both buffers */
    int neg = 0;
    for (int q = 0; q < Q; q++) {
        double feq = calcfeq(o, CX[q], CY[q], CZ[q], o->w[q], 0.0, 0.0, 0.0, 1.0, 0.0, &neg);
        double *a = o->f1 + (size_t)q * o->N3, *b = o->f2 + (size_t)q * o->N3;
#pragma omp parallel for
        for (size_t n = 0; n < o->N3; n++) { a[n] = feq; b[n] = feq; }
    }
    return o;
}

void oracle3d_destroy(void *h) {
    oracle3d *o = (oracle3d *)h;
    if (!o) return;
    free(o->f1); free(o->f2); free(o);
}

static inline int inside(int a, int n) { return a >= 0 && a < n; }

/* The collision of one cell, 3D:224-246 (BGK) / 3D:268-302 (TRT): f[19] (the reference's f1[x][:]) ->
 * fs[19] (the values macrocollideandstream pushes to the neighbours; fs[0] stays local). */
static inline void collide_cell(const oracle3d *o, const double *f, double *fs, int *neg) {
    if (o->p.METHOD == 0) {
        double rho = 0.0, u = 0.0, v = 0.0, w = 0.0;
        for (int q = 0; q < Q; q++) {
            double ft = f[q];
            rho += ft;
            u += CX[q] * ft;
            v += CY[q] * ft;
            w += CZ[q] * ft;
        }
        u /= rho; v /= rho; w /= rho;
        double U2 = (u * u) + (v * v) + (w * w);
        for (int q = 0; q < Q; q++) {
            double feq = calcfeq(o, CX[q], CY[q], CZ[q], o->w[q], u, v, w, rho, U2, neg);
            fs[q] = o->OmOMEGA * f[q] + o->OMEGA * feq;
        }
    } else {
        double rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8] + f[9] +
                     f[10] + f[11] + f[12] + f[13] + f[14] + f[15] + f[16] + f[17] + f[18];
        double rinv = 1.0 / rho;
        double u = (CX[1] * f[1] + CX[2] * f[2] + CX[7] * f[7] + CX[8] * f[8] + CX[9] * f[9] +
                    CX[10] * f[10] + CX[13] * f[13] + CX[14] * f[14] + CX[15] * f[15] +
                    CX[16] * f[16]) * rinv;
        double v = (CY[3] * f[3] + CY[4] * f[4] + CY[7] * f[7] + CY[8] * f[8] + CY[11] * f[11] +
                    CY[12] * f[12] + CY[13] * f[13] + CY[14] * f[14] + CY[17] * f[17] +
                    CY[18] * f[18]) * rinv;
        double w = (CZ[5] * f[5] + CZ[6] * f[6] + CZ[9] * f[9] + CZ[10] * f[10] + CZ[11] * f[11] +
                    CZ[12] * f[12] + CZ[15] * f[15] + CZ[16] * f[16] + CZ[17] * f[17] +
                    CZ[18] * f[18]) * rinv;
        double U2 = (u * u) + (v * v) + (w * w);
        fs[0] = f[0] - o->OMEGA * (f[0] - calcfeq(o, 0, 0, 0, o->w[0], u, v, w, rho, U2, neg));
        for (int q = 1; q < Q; q += 2) {
            int op = q + 1;
            double feq1 = calcfeq(o, CX[q], CY[q], CZ[q], o->w[q], u, v, w, rho, U2, neg);
            double feq2 = calcfeq(o, CX[op], CY[op], CZ[op], o->w[op], u, v, w, rho, U2, neg);
            double fplus = 0.5 * (f[q] + f[op]);
            double fminus = 0.5 * (f[q] - f[op]);
            double feqplus = 0.5 * (feq1 + feq2);
            double feqminus = 0.5 * (feq1 - feq2);
            fplus = o->OMEGA * (fplus - feqplus);
            fminus = o->OMEGAm * (fminus - feqminus);
            fs[q] = f[q] - fplus - fminus;
            fs[op] = f[op] - fplus + fminus;
        }
    }
}

/* macrocollideandstream, 3D:215-321: collide, then push to x + c_q where that is inside the box
 * (3D:252, 308, 314); the rest population is written locally (3D:286 / q = 0 in 3D:249-253). */
static void collide_and_stream(oracle3d *o) {
    const int N = o->p.N;
    const double *f1 = o->f1;
    double *f2 = o->f2;
    int neg_any = 0;
#pragma omp parallel for collapse(2) reduction(| : neg_any)
    for (int k = 0; k < N; k++) {
        for (int j = 0; j < N; j++) {
            for (int i = 0; i < N; i++) {
                double f[Q], fs[Q];
                int neg = 0;
                for (int q = 0; q < Q; q++) f[q] = f1[LU4(o, i, j, k, q)];
                collide_cell(o, f, fs, &neg);
                for (int q = 0; q < Q; q++) {
                    int in = i + CX[q], jn = j + CY[q], kn = k + CZ[q];
                    if (inside(in, N) && inside(jn, N) && inside(kn, N)) f2[LU4(o, in, jn, kn, q)] = fs[q];
                }
                neg_any |= neg;
            }
        }
    }
    if (neg_any) o->unstable = 1;
}

/* ---- HechtBC, 3D:385-675, table-driven restatement (cell-local) -------------- */

/* One wall face (not the lid): normal unknown `un` <- `on`; two tangential corrections
 * Na = 0.5*((p0+p1+p2) - (m0+m1+m2)); four diagonal unknowns  u <- o + sgn * N[n]. */
typedef struct {
    int un, on;
    int p[2][3], m[2][3];
    struct { int u, o, sgn, n; } a[4];
} face_rule;

static const face_rule FACE_Z0 = {5, 6, {{1, 7, 13}, {3, 7, 14}}, {{2, 14, 8}, {4, 13, 8}},     /* 3D:392-398 */
                                  {{9, 10, -1, 0}, {16, 15, +1, 0}, {11, 12, -1, 1}, {18, 17, +1, 1}}};
static const face_rule FACE_ZM = {6, 5, {{1, 7, 13}, {3, 7, 14}}, {{2, 14, 8}, {4, 13, 8}},     /* 3D:401-407 */
                                  {{10, 9, +1, 0}, {15, 16, -1, 0}, {12, 11, +1, 1}, {17, 18, -1, 1}}};
static const face_rule FACE_X0 = {1, 2, {{3, 11, 17}, {5, 14, 11}}, {{4, 18, 12}, {6, 17, 12}}, /* 3D:410-416 */
                                  {{13, 14, +1, 0}, {7, 8, -1, 0}, {9, 10, -1, 1}, {15, 16, +1, 1}}};
static const face_rule FACE_XM = {2, 1, {{3, 11, 17}, {5, 14, 11}}, {{4, 18, 12}, {6, 17, 12}}, /* 3D:419-425 */
                                  {{14, 13, -1, 0}, {8, 7, +1, 0}, {10, 9, +1, 1}, {16, 15, -1, 1}}};
static const face_rule FACE_Y0 = {3, 4, {{1, 9, 15}, {5, 9, 16}}, {{2, 16, 10}, {6, 15, 10}},   /* 3D:428-434 */
                                  {{7, 8, -1, 0}, {14, 13, +1, 0}, {11, 12, -1, 1}, {17, 18, +1, 1}}};

static inline void apply_face(const oracle3d *o, double *F, size_t cell, const face_rule *r) {
    const size_t S = o->N3;
#define G(q) F[(size_t)(q) * S + cell]
    G(r->un) = G(r->on);
    double Nn[2];
    for (int n = 0; n < 2; n++)
        Nn[n] = 0.5 * ((G(r->p[n][0]) + G(r->p[n][1]) + G(r->p[n][2])) -
                       (G(r->m[n][0]) + G(r->m[n][1]) + G(r->m[n][2])));
    for (int a = 0; a < 4; a++) {
        if (r->a[a].sgn > 0) G(r->a[a].u) = G(r->a[a].o) + Nn[r->a[a].n];
        else                 G(r->a[a].u) = G(r->a[a].o) - Nn[r->a[a].n];
    }
#undef G
}

/* the moving lid y = N-1, 3D:437-445 */
static inline void apply_lid(const oracle3d *o, double *F, size_t cell, double vfrac) {
    const size_t S = o->N3;
    const double ulid = o->ulid, vlid = o->vlid, wlid = o->wlid, third = o->third, sixth = o->sixth;
#define G(q) F[(size_t)(q) * S + cell]
    double rho = vfrac * (G(1) + G(2) + G(5) + G(6) + G(9) + G(15) + G(16) + G(10) + G(0) +
                          2.0 * (G(3) + G(7) + G(14) + G(11) + G(17)));
    G(4) = G(3) - rho * vlid * third;
    double Nyx = 0.5 * ((G(1) + G(9) + G(15)) - (G(2) + G(16) + G(10))) - rho * ulid * third;
    double Nyz = 0.5 * ((G(5) + G(9) + G(16)) - (G(6) + G(15) + G(10))) - rho * wlid * third;
    G(8) = G(7) - rho * (vlid + ulid) * sixth + Nyx;
    G(13) = G(14) + rho * (-vlid + ulid) * sixth - Nyx;
    G(12) = G(11) - rho * (vlid + wlid) * sixth + Nyz;
    G(18) = G(17) + rho * (-vlid + wlid) * sixth - Nyz;
#undef G
}

/* One edge: temp = coef * (f[ta] - f[tb]); nine unknowns  u <- opp(u) + sgn * temp, in source
 * order.  bx/by/bz: 0 = running index, 1 = coordinate 0, 2 = coordinate N-1. */
typedef struct {
    int bx, by, bz;
    double coef;
    int ta, tb;
    struct { int u, sgn; } a[9];
} edge_rule;

static const edge_rule EDGES[12] = {
    {0, 1, 1, 0.25, 1, 2, {{3, 0}, {7, -1}, {11, 0}, {14, +1}, {17, 0}, {5, 0}, {9, -1}, {16, +1}, {18, 0}}},   /* 3D:452-463 */
    {0, 1, 2, 0.25, 1, 2, {{3, 0}, {7, -1}, {11, 0}, {14, +1}, {17, 0}, {6, 0}, {10, +1}, {12, 0}, {15, -1}}},  /* 3D:466-476 */
    {0, 2, 1, 0.25, 1, 2, {{4, 0}, {8, +1}, {12, 0}, {13, -1}, {18, 0}, {5, 0}, {9, -1}, {11, 0}, {16, +1}}},   /* 3D:479-489 */
    {0, 2, 2, 0.25, 1, 2, {{4, 0}, {8, +1}, {12, 0}, {13, -1}, {18, 0}, {6, 0}, {10, +1}, {15, -1}, {17, 0}}},  /* 3D:492-502 */
    {1, 0, 1, 0.25, 3, 4, {{1, 0}, {7, -1}, {9, 0}, {13, +1}, {15, 0}, {5, 0}, {11, -1}, {16, 0}, {18, +1}}},   /* 3D:505-515 */
    {1, 0, 2, 0.25, 3, 4, {{1, 0}, {7, -1}, {9, 0}, {13, +1}, {15, 0}, {6, 0}, {10, 0}, {12, +1}, {17, -1}}},   /* 3D:518-528 */
    {2, 0, 1, 0.25, 3, 4, {{2, 0}, {8, -1}, {10, 0}, {14, +1}, {16, 0}, {5, 0}, {9, 0}, {11, -1}, {18, +1}}},   /* 3D:531-541 */
    {2, 0, 2, -0.25, 3, 4, {{2, 0}, {8, -1}, {10, 0}, {14, +1}, {16, 0}, {6, 0}, {12, -1}, {15, 0}, {17, +1}}}, /* 3D:544-554 */
    {1, 1, 0, 0.25, 5, 6, {{3, 0}, {7, 0}, {11, -1}, {14, 0}, {17, +1}, {1, 0}, {9, -1}, {13, 0}, {15, +1}}},   /* 3D:557-567 */
    {1, 2, 0, 0.25, 5, 6, {{4, 0}, {8, 0}, {12, +1}, {13, 0}, {18, -1}, {1, 0}, {7, 0}, {9, -1}, {15, +1}}},    /* 3D:570-580 */
    {2, 1, 0, 0.25, 5, 6, {{3, 0}, {7, 0}, {11, -1}, {14, 0}, {17, +1}, {2, 0}, {8, 0}, {10, +1}, {16, -1}}},   /* 3D:584-594 */
    {2, 2, 0, 0.25, 5, 6, {{4, 0}, {8, 0}, {12, +1}, {13, 0}, {18, -1}, {2, 0}, {10, +1}, {14, 0}, {16, -1}}},  /* 3D:598-608 */
};

static inline int opp(int q) { return q == 0 ? 0 : ((q & 1) ? q + 1 : q - 1); }

static inline void apply_edge(const oracle3d *o, double *F, size_t cell, const edge_rule *r) {
    const size_t S = o->N3;
#define G(q) F[(size_t)(q) * S + cell]
    double temp = r->coef * (G(r->ta) - G(r->tb));
    for (int a = 0; a < 9; a++) {
        int u = r->a[a].u, s = r->a[a].sgn;
        if (s == 0)     G(u) = G(opp(u));
        else if (s > 0) G(u) = G(opp(u)) + temp;
        else            G(u) = G(opp(u)) - temp;
    }
#undef G
}

/* Corners, 3D:611-674: six unknowns u <- opp(u); (bx,by,bz) in {1 = 0, 2 = N-1}. */
typedef struct { int bx, by, bz; int u[6]; } corner_rule;
static const corner_rule CORNERS[8] = {
    {1, 2, 1, {1, 4, 5, 9, 13, 18}},   /* 3D:613-618 */
    {1, 1, 1, {1, 3, 5, 9, 7, 11}},    /* 3D:621-626 */
    {2, 2, 1, {2, 4, 5, 16, 8, 18}},   /* 3D:629-634 */
    {2, 1, 1, {2, 3, 5, 16, 14, 11}},  /* 3D:637-642 */
    {1, 2, 2, {1, 4, 6, 15, 13, 12}},  /* 3D:645-650 */
    {1, 1, 2, {1, 3, 6, 15, 7, 17}},   /* 3D:653-658 */
    {2, 2, 2, {2, 4, 6, 10, 8, 12}},   /* 3D:661-666 */
    {2, 1, 2, {2, 3, 6, 10, 14, 17}},  /* 3D:669-674 */
};

static inline int coord(int b, int run, int M) { return b == 0 ? run : (b == 1 ? 0 : M); }

static void hecht_bc(oracle3d *o, double *F) {
    const int N = o->p.N, M = N - 1;
    const double vfrac = 1.0 / (o->vlid + 1.0); /* 3D:386 */
#define CELL(i, j, k) ((size_t)(k) * o->N2 + (size_t)(j) * o->N1 + (size_t)(i))
#pragma omp parallel for
    for (int a = 1; a < M; a++) {
        for (int b = 1; b < M; b++) {
            /* the reference's (i,j) pair is used as (x,y), (y,z) or (x,z) per face */
            apply_face(o, F, CELL(a, b, 0), &FACE_Z0);
            apply_face(o, F, CELL(a, b, M), &FACE_ZM);
            apply_face(o, F, CELL(0, a, b), &FACE_X0);
            apply_face(o, F, CELL(M, a, b), &FACE_XM);
            apply_face(o, F, CELL(a, 0, b), &FACE_Y0);
            apply_lid(o, F, CELL(a, M, b), vfrac);
        }
    }
#pragma omp parallel for
    for (int r = 1; r < M; r++)
        for (int e = 0; e < 12; e++) {
            const edge_rule *er = &EDGES[e];
            apply_edge(o, F, CELL(coord(er->bx, r, M), coord(er->by, r, M), coord(er->bz, r, M)), er);
        }
    for (int c = 0; c < 8; c++) {
        const corner_rule *cr = &CORNERS[c];
        size_t cell = CELL(coord(cr->bx, 0, M), coord(cr->by, 0, M), coord(cr->bz, 0, M));
        for (int a = 0; a < 6; a++)
            F[(size_t)cr->u[a] * o->N3 + cell] = F[(size_t)opp(cr->u[a]) * o->N3 + cell];
    }
#undef CELL
}

/* diff, 3D:678-685 (summation order is unspecified in the reference: OpenMP reduction) */
static double l2diff(const double *a, const double *b, size_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (size_t i = 0; i < n; i++) s += (a[i] - b[i]) * (a[i] - b[i]);
    return sqrt(s);
}

/* n iterations of: macrocollideandstream(); HechtBC(f2); [diff]; swap  (3D:94-100, 106-123).
 * Returns 0, or 1 if a negative equilibrium was seen (the reference would have exit(1)'d;
 * here the step completes and stepping stops). */
int oracle3d_step(void *h, int n, int nthreads, int want_diff) {
    oracle3d *o = (oracle3d *)h;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int t = 0; t < n; t++) {
        if (o->unstable) return 1;
        collide_and_stream(o);
        hecht_bc(o, o->f2);
        if (want_diff && t == n - 1) o->last_diff = l2diff(o->f2, o->f1, o->NQ);
        double *tmp = o->f1; o->f1 = o->f2; o->f2 = tmp; /* swap, 3D:722-727 */
        o->steps_done++;
    }
    return o->unstable;
}

double oracle3d_last_diff(void *h) { return ((oracle3d *)h)->last_diff; }
int oracle3d_unstable(void *h) { return ((oracle3d *)h)->unstable; }

/* newest lattice, NQ doubles in LU4 order (after the swap that is f1) */
void oracle3d_get_f(void *h, double *out) {
    oracle3d *o = (oracle3d *)h;
    memcpy(out, o->f1, o->NQ * sizeof(double));
}

/* the buffer that will be overwritten next (two steps old where streaming does not reach) */
void oracle3d_get_f_prev(void *h, double *out) {
    oracle3d *o = (oracle3d *)h;
    memcpy(out, o->f2, o->NQ * sizeof(double));
}

void oracle3d_derived(void *h, double *out8) {
    oracle3d *o = (oracle3d *)h;
    out8[0] = o->NU; out8[1] = o->TAU; out8[2] = o->OMEGA; out8[3] = o->OMEGAm;
    out8[4] = o->OmOMEGA; out8[5] = o->ulid; out8[6] = o->vlid; out8[7] = o->wlid;
}

/* macrovars (3D:695-719) + calcvmag (3D:730-738) on the newest lattice; LU3 order (3D:47) */
void oracle3d_macrovars(void *h, double *u1, double *u2, double *u3, double *rho, double *vmag) {
    oracle3d *o = (oracle3d *)h;
    const double *F = o->f1;
#pragma omp parallel for
    for (size_t c = 0; c < o->N3; c++) {
        double r = 0.0, u = 0.0, v = 0.0, w = 0.0;
        for (int q = 0; q < Q; q++) {
            double ft = F[(size_t)q * o->N3 + c];
            r += ft;
            u += CX[q] * ft;
            v += CY[q] * ft;
            w += CZ[q] * ft;
        }
        rho[c] = r;
        u1[c] = u / r;
        u2[c] = v / r;
        u3[c] = w / r;
        vmag[c] = sqrt((u1[c] * u1[c]) + (u2[c] * u2[c]) + (u3[c] * u3[c]));
    }
}

/* Timed loop for bench.py (cpu_baseline kind "port"): n steps, returns seconds. */
double oracle3d_time_steps(void *h, int n, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = omp_get_wtime();
    oracle3d_step(h, n, nthreads, 0);
    return omp_get_wtime() - t0;
#else
    oracle3d_step(h, n, nthreads, 0);
    return 0.0;
#endif
}

/* ---- single-cell entry points for tests of the product's device functions compiled for the host
 *      (tests/host_device_functions_check.cu) ------------------------------------------------------ */
int oracle3d_collide_cell(void *h, const double *f19, double *out19) {
    int neg = 0;
    collide_cell((const oracle3d *)h, f19, out19, &neg);
    return neg;
}

/* HechtBC rule of the boundary class (bx, by, bz) in {0: inside, 1: coordinate 0, 2: coordinate N-1}
 * applied in place to ONE cell's 19 populations.  Returns 0 if the class has no rule (interior). */
int oracle3d_rule_cell(void *h, int bx, int by, int bz, double *f19) {
    oracle3d tmp = *(const oracle3d *)h;
    tmp.N3 = 1; /* population stride of a single cell */
    const int nb = (bx != 0) + (by != 0) + (bz != 0);
    if (nb == 0) return 0;
    if (nb == 1) {
        if (bz == 1) apply_face(&tmp, f19, 0, &FACE_Z0);
        else if (bz == 2) apply_face(&tmp, f19, 0, &FACE_ZM);
        else if (bx == 1) apply_face(&tmp, f19, 0, &FACE_X0);
        else if (bx == 2) apply_face(&tmp, f19, 0, &FACE_XM);
        else if (by == 1) apply_face(&tmp, f19, 0, &FACE_Y0);
        else apply_lid(&tmp, f19, 0, 1.0 / (tmp.vlid + 1.0));
        return 1;
    }
    if (nb == 2) {
        for (int e = 0; e < 12; e++)
            if (EDGES[e].bx == bx && EDGES[e].by == by && EDGES[e].bz == bz) { apply_edge(&tmp, f19, 0, &EDGES[e]); return 2; }
        return -1;
    }
    for (int c = 0; c < 8; c++)
        if (CORNERS[c].bx == bx && CORNERS[c].by == by && CORNERS[c].bz == bz) {
            for (int a = 0; a < 6; a++) f19[CORNERS[c].u[a]] = f19[opp(CORNERS[c].u[a])];
            return 3;
        }
    return -1;
}
