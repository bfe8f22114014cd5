/*
 * oracle/lbm2d_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle of the D2Q9 variant, config 5).
 *
 * Restates the cell update of the reference's 2D program
 *   /root/reference/2D Lid-driven Cavity with Lattice Boltzmann Method/main.cpp   ("2D:")
 * streamCollide 2D:281-416, macroVars 2D:419-449, convergence 2D:452-478, constants 2D:12-31.
 *
 * TWO time-stepping schemes over the SAME cell arithmetic (functions pull and collide below):
 *
 *  (A) oracle2d_*  "synchronous": two lattices, every cell reads step t and writes step t+1.
 *      This is what the GPU kernel computes and the GPU is tested bit-for-bit
 *      against it.  PARITY NOTE: this scheme is NOT what the reference executes, so it is
 *      "parity unpinned by the reference itself" at the per-iteration level; it is anchored to
 *      the reference through (B): same cell arithmetic, same fixed point.
 *
 *  (B) inplace2d_*  a literal emulation of what the reference executes: ONE lattice per MPI rank,
 *      updated in place in sweep order (i outer, j inner; interior rows, then the top row, then
 *      the bottom row, 2D:166-178), ghost rows filled at the START of each iteration with the
 *      neighbour's values from the END of the previous one (2D:162-163, 208-270), P >= 1 ranks
 *      emulated serially (ranks are independent within an iteration).  PINNED: bit-for-bit
 *      against the real reference built over a fork/shared-memory MPI shim (oracle/mpi_shim/,
 *      oracle/build_ref.py -> oracle/_ref/lbm2d_ref_*), see tests/test_oracle2d.py, and against
 *      the steady-state values SURVEY App. D records from that same program.
 *
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q 9

typedef struct {
    int N, M;
    double Re, Ulid, MAGIC;
} oracle2d_params;

typedef struct {
    double NU, TAU, OMEGA, OMEGAm, w0, wS, wd, lid6;
} consts2d;

static void derive(const oracle2d_params *p, consts2d *c) {
    c->NU = p->Ulid * p->N / p->Re;                       /* 2D:22 */
    c->TAU = c->NU * 3.0 + 0.5;                           /* 2D:23 */
    c->OMEGA = 1.0 / c->TAU;                              /* 2D:24 */
    c->OMEGAm = 1.0 / (0.5 + (p->MAGIC / (c->TAU - 0.5))); /* 2D:26 */
    c->w0 = 4.0 / 9.0; c->wS = 1.0 / 9.0; c->wd = 1.0 / 36.0; /* 2D:27-29 */
    c->lid6 = 6.0 * c->wd * p->Ulid;                      /* 2D:354 */
}

/* Collision of 2D:360-413: ft[9] (streamed, pre-collision) -> out[9]. */
static inline void collide(const consts2d *c, const double *ft, double *out) {
    const double OMEGA = c->OMEGA, OMEGAm = c->OMEGAm, w0 = c->w0, wS = c->wS, wd = c->wd;
    const double ft0 = ft[0], ft1 = ft[1], ft2 = ft[2], ft3 = ft[3], ft4 = ft[4], ft5 = ft[5], ft6 = ft[6],
                 ft7 = ft[7], ft8 = ft[8];
    const double drho = ft0 + ft1 + ft2 + ft3 + ft4 + ft5 + ft6 + ft7 + ft8;
    const double ux = ft1 + ft5 + ft8 - (ft3 + ft6 + ft7);
    const double uy = ft2 + ft5 + ft6 - (ft4 + ft7 + ft8);
    double fplus, fminus, feqTerm1, feqTerm2, cdotu;
    const double halfOmega = 0.5 * OMEGA;
    const double halfOmegaMinus = 0.5 * OMEGAm;
    const double u2x15 = 1.5 * ((ux * ux) + (uy * uy));
    const double wsu2x15 = wS * u2x15;
    const double wdu2x15 = wd * u2x15;
    const double wsdrho = wS * drho;
    const double wddrho = wd * drho;
    out[0] = ft0 - OMEGA * (ft0 - w0 * (drho - u2x15));
    cdotu = ux;
    feqTerm1 = wsdrho + wS * (4.5 * (cdotu * cdotu)) - wsu2x15;
    feqTerm2 = 3.0 * wS * cdotu;
    fplus = halfOmega * ((ft1 + ft3) - 2 * feqTerm1);
    fminus = halfOmegaMinus * ((ft1 - ft3) - 2 * feqTerm2);
    out[1] = ft1 - fplus - fminus;
    out[3] = ft3 - fplus + fminus;
    cdotu = uy;
    feqTerm1 = wsdrho + wS * (4.5 * (cdotu * cdotu)) - wsu2x15;
    feqTerm2 = 3.0 * wS * cdotu;
    fplus = halfOmega * ((ft2 + ft4) - 2 * feqTerm1);
    fminus = halfOmegaMinus * ((ft2 - ft4) - 2 * feqTerm2);
    out[2] = ft2 - fplus - fminus;
    out[4] = ft4 - fplus + fminus;
    cdotu = ux + uy;
    feqTerm1 = wddrho + wd * (4.5 * (cdotu * cdotu)) - wdu2x15;
    feqTerm2 = 3.0 * wd * cdotu;
    fplus = halfOmega * ((ft5 + ft7) - 2 * feqTerm1);
    fminus = halfOmegaMinus * ((ft5 - ft7) - 2 * feqTerm2);
    out[5] = ft5 - fplus - fminus;
    out[7] = ft7 - fplus + fminus;
    cdotu = -ux + uy;
    feqTerm1 = wddrho + wd * (4.5 * (cdotu * cdotu)) - wdu2x15;
    feqTerm2 = 3.0 * wd * cdotu;
    fplus = halfOmega * ((ft6 + ft8) - 2 * feqTerm1);
    fminus = halfOmegaMinus * ((ft6 - ft8) - 2 * feqTerm2);
    out[6] = ft6 - fplus - fminus;
    out[8] = ft8 - fplus + fminus;
}

/* The pull of 2D:289-356 for the cell (i, j) of an array `d` in the reference's layout
 * LU3(i,j,k) = 9*N*j + N*k + i (2D:276-278) whose rows [0, jEndWGhosts) exist; `bottom`/`top`
 * say whether row j = 0 is the wall row / the last row is the lid row (2D:85-94). */
#define LU3(i, j, k) ((size_t)Q * N * (j) + (size_t)N * (k) + (i))
static inline void pull(const consts2d *c, const double *d, int N, int i, int j, int jEndWGhosts, int bottom, int top,
                        double *ft) {
    const int im1 = i - 1, ip1 = i + 1, jm1 = j - 1, jp1 = j + 1;
    double ft0, ft1 = 0, ft2 = 0, ft3 = 0, ft4 = 0, ft5 = 0, ft6 = 0, ft7 = 0, ft8 = 0;
    ft0 = d[LU3(i, j, 0)];
    if (i > 0) {
        ft1 = d[LU3(im1, j, 1)];
        if (j > 0) ft5 = d[LU3(im1, jm1, 5)];
        if (jp1 < jEndWGhosts) ft8 = d[LU3(im1, jp1, 8)];
    } else { /* left wall */
        ft1 = d[LU3(i, j, 3)];
        ft5 = d[LU3(i, j, 7)];
        ft8 = d[LU3(i, j, 6)];
    }
    if (ip1 < N) {
        ft3 = d[LU3(ip1, j, 3)];
        if (j > 0) ft6 = d[LU3(ip1, jm1, 6)];
        if (jp1 < jEndWGhosts) ft7 = d[LU3(ip1, jp1, 7)];
    } else { /* right wall */
        ft3 = d[LU3(i, j, 1)];
        ft6 = d[LU3(i, j, 8)];
        ft7 = d[LU3(i, j, 5)];
    }
    if (j > 0) {
        ft2 = d[LU3(i, jm1, 2)];
    } else if (bottom) {
        ft2 = d[LU3(i, j, 4)];
        ft5 = d[LU3(i, j, 7)];
        ft6 = d[LU3(i, j, 8)];
    }
    if (jp1 < jEndWGhosts) {
        ft4 = d[LU3(i, jp1, 4)];
    } else if (top) {
        ft4 = d[LU3(i, j, 2)];
        ft7 = d[LU3(i, j, 5)] - c->lid6;
        ft8 = d[LU3(i, j, 6)] + c->lid6;
    }
    ft[0] = ft0; ft[1] = ft1; ft[2] = ft2; ft[3] = ft3; ft[4] = ft4; ft[5] = ft5; ft[6] = ft6; ft[7] = ft7; ft[8] = ft8;
}

/* ===== (A) synchronous two-lattice scheme ==================================================== */
typedef struct {
    oracle2d_params p;
    consts2d c;
    double *a, *b;          /* reference layout, M rows, no ghosts */
    double *u, *v, *rho;    /* macroVars state */
    long steps;
} sync2d;

void *oracle2d_create(const oracle2d_params *p) {
    sync2d *s = (sync2d *)calloc(1, sizeof(sync2d));
    s->p = *p;
    derive(p, &s->c);
    const size_t n = (size_t)Q * p->N * p->M, m = (size_t)p->N * p->M;
    s->a = (double *)calloc(n, sizeof(double));   /* dist = 0, 2D:127 */
    s->b = (double *)calloc(n, sizeof(double));
    s->u = (double *)calloc(m, sizeof(double));
    s->v = (double *)calloc(m, sizeof(double));
    s->rho = (double *)malloc(m * sizeof(double));
    for (size_t k = 0; k < m; k++) s->rho[k] = 1.0;
    return s;
}

void oracle2d_destroy(void *h) {
    sync2d *s = (sync2d *)h;
    if (!s) return;
    free(s->a); free(s->b); free(s->u); free(s->v); free(s->rho); free(s);
}

void oracle2d_step(void *h, int n) {
    sync2d *s = (sync2d *)h;
    const int N = s->p.N, M = s->p.M;
    for (int t = 0; t < n; t++) {
#pragma omp parallel for
        for (int j = 0; j < M; j++)
            for (int i = 0; i < N; i++) {
                double ft[Q], out[Q];
                pull(&s->c, s->a, N, i, j, M, 1, 1, ft);
                collide(&s->c, ft, out);
                for (int k = 0; k < Q; k++) s->b[LU3(i, j, k)] = out[k];
            }
        double *tmp = s->a; s->a = s->b; s->b = tmp;
        s->steps++;
    }
}

/* macroVars (2D:419-449) on `d`; returns the accumulated diff (not yet sqrt'ed). */
static double macro_diff(const double *d, int N, int jStart, int jEnd, int rowoff, double *u, double *v, double *rho) {
    double diff = 0.0;
    for (int i = 0; i < N; i++)
        for (int j = jStart; j < jEnd; j++) {
            const size_t ind1 = (size_t)N * (j - rowoff) + i;
            const double f0 = d[LU3(i, j, 0)], f1 = d[LU3(i, j, 1)], f2 = d[LU3(i, j, 2)], f3 = d[LU3(i, j, 3)],
                         f4 = d[LU3(i, j, 4)], f5 = d[LU3(i, j, 5)], f6 = d[LU3(i, j, 6)], f7 = d[LU3(i, j, 7)],
                         f8 = d[LU3(i, j, 8)];
            const double new_r = 1.0 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 + f8;
            const double new_u = f1 - f3 + f5 - f6 - f7 + f8;
            const double new_v = f2 - f4 + f5 + f6 - f7 - f8;
            diff += pow(u[ind1] - new_u, 2) + pow(v[ind1] - new_v, 2) + pow(rho[ind1] - new_r, 2);
            u[ind1] = new_u;
            rho[ind1] = new_r;
            v[ind1] = new_v;
        }
    return diff;
}

/* sqrt(sum (du^2 + dv^2 + drho^2)) against the previous call; updates u, v, rho (2D:419-460). */
double oracle2d_residual(void *h) {
    sync2d *s = (sync2d *)h;
    return sqrt(macro_diff(s->a, s->p.N, 0, s->p.M, 0, s->u, s->v, s->rho));
}

void oracle2d_get_dist(void *h, double *out) { /* reference layout, 9*N*M doubles */
    sync2d *s = (sync2d *)h;
    memcpy(out, s->a, (size_t)Q * s->p.N * s->p.M * sizeof(double));
}

void oracle2d_get_macro(void *h, double *u, double *v, double *rho) {
    sync2d *s = (sync2d *)h;
    const size_t m = (size_t)s->p.N * s->p.M;
    memcpy(u, s->u, m * sizeof(double)); memcpy(v, s->v, m * sizeof(double)); memcpy(rho, s->rho, m * sizeof(double));
}

void oracle2d_derived(const oracle2d_params *p, double *out4) {
    consts2d c;
    derive(p, &c);
    out4[0] = c.NU; out4[1] = c.TAU; out4[2] = c.OMEGA; out4[3] = c.OMEGAm;
}

/* ===== (B) literal in-place emulation with P ranks =========================================== */
typedef struct {
    int bottom, top, Mj, rowStart;               /* 2D:85-105 */
    int jStartWoGhosts, jEndWoGhosts, jEndWGhosts; /* 2D:107-125 */
    double *dist, *u, *v, *rho;
} rank2d;

typedef struct {
    oracle2d_params p;
    consts2d c;
    int P;
    rank2d *r;
    long steps;
} inplace2d;

void *inplace2d_create(const oracle2d_params *p, int P) {
    inplace2d *s = (inplace2d *)calloc(1, sizeof(inplace2d));
    s->p = *p; s->P = P;
    derive(p, &s->c);
    s->r = (rank2d *)calloc((size_t)P, sizeof(rank2d));
    const int N = p->N, M = p->M;
    for (int rank = 0; rank < P; rank++) {
        rank2d *r = &s->r[rank];
        /* region flags 2D:85-94: rank 0 is BOTTOM, rank P-1 is TOP, else INTERNAL (with P == 1 the
         * reference refuses to run, 2D:69-72; the emulation then treats the single rank as both) */
        r->bottom = (rank == 0);
        r->top = (rank == P - 1);
        if (rank < M % P) { r->Mj = M / P + 1; r->rowStart = rank * r->Mj; }   /* 2D:98-104 */
        else              { r->Mj = M / P;     r->rowStart = M - (P - rank) * r->Mj; }
        int rows;
        if (r->bottom && r->top) { rows = r->Mj; r->jStartWoGhosts = 0; r->jEndWoGhosts = r->Mj; r->jEndWGhosts = r->Mj; }
        else if (!r->bottom && !r->top) { rows = r->Mj + 2; r->jStartWoGhosts = 1; r->jEndWoGhosts = r->Mj + 1; r->jEndWGhosts = r->Mj + 2; }
        else if (r->top) { rows = r->Mj + 1; r->jStartWoGhosts = 1; r->jEndWoGhosts = r->Mj + 1; r->jEndWGhosts = r->Mj + 1; }
        else { rows = r->Mj + 1; r->jStartWoGhosts = 0; r->jEndWoGhosts = r->Mj; r->jEndWGhosts = r->Mj + 1; }
        r->dist = (double *)calloc((size_t)Q * N * rows, sizeof(double));
        r->u = (double *)calloc((size_t)N * r->Mj, sizeof(double));
        r->v = (double *)calloc((size_t)N * r->Mj, sizeof(double));
        r->rho = (double *)malloc((size_t)N * r->Mj * sizeof(double));
        for (size_t k = 0; k < (size_t)N * r->Mj; k++) r->rho[k] = 1.0;
    }
    return s;
}

void inplace2d_destroy(void *h) {
    inplace2d *s = (inplace2d *)h;
    if (!s) return;
    for (int k = 0; k < s->P; k++) { free(s->r[k].dist); free(s->r[k].u); free(s->r[k].v); free(s->r[k].rho); }
    free(s->r); free(s);
}

/* streamCollide(dist, yStart, yEnd), 2D:281-416: in place, i outer, j inner */
static void stream_collide_inplace(const inplace2d *s, rank2d *r, int yStart, int yEnd) {
    const int N = s->p.N;
    for (int i = 0; i < N; i++)
        for (int j = yStart; j < yEnd; j++) {
            double ft[Q], out[Q];
            pull(&s->c, r->dist, N, i, j, r->jEndWGhosts, r->bottom, r->top, ft);
            collide(&s->c, ft, out);
            for (int k = 0; k < Q; k++) r->dist[LU3(i, j, k)] = out[k];
        }
}

void inplace2d_step(void *h, int n) {
    inplace2d *s = (inplace2d *)h;
    const int N = s->p.N, P = s->P;
    static const int UPQ[3] = {2, 5, 6}, DNQ[3] = {4, 7, 8};
    for (int t = 0; t < n; t++) {
        /* exchangeDataDownward / exchangeDataUpward (2D:208-270): posted at the START of the
         * iteration, i.e. ghost rows receive the neighbour's rows as they are now */
        for (int rank = 0; rank < P; rank++) {
            rank2d *r = &s->r[rank];
            if (!r->top) { /* my top row goes up (k = 2,5,6) into the ghost row below rank+1's first row */
                rank2d *ab = &s->r[rank + 1];
                for (int a = 0; a < 3; a++)
                    memcpy(&ab->dist[LU3(0, ab->jStartWoGhosts - 1, UPQ[a])], &r->dist[LU3(0, r->jEndWoGhosts - 1, UPQ[a])],
                           (size_t)N * sizeof(double));
            }
            if (!r->bottom) { /* my bottom row goes down (k = 4,7,8) into rank-1's top ghost row */
                rank2d *be = &s->r[rank - 1];
                for (int a = 0; a < 3; a++)
                    memcpy(&be->dist[LU3(0, be->jEndWoGhosts, DNQ[a])], &r->dist[LU3(0, r->jStartWoGhosts, DNQ[a])],
                           (size_t)N * sizeof(double));
            }
        }
        /* 2D:166-178: interior rows, then the row touching the top ghosts, then the bottom one */
#pragma omp parallel for
        for (int rank = 0; rank < P; rank++) {
            rank2d *r = &s->r[rank];
            stream_collide_inplace(s, r, r->jStartWoGhosts + 1, r->jEndWoGhosts - 1);
            stream_collide_inplace(s, r, r->jEndWoGhosts - 1, r->jEndWoGhosts);
            stream_collide_inplace(s, r, r->jStartWoGhosts, r->jStartWoGhosts + 1);
        }
        s->steps++;
    }
}

double inplace2d_residual(void *h) { /* macroVars + Allreduce + sqrt, 2D:419-460 (sum in rank order) */
    inplace2d *s = (inplace2d *)h;
    double tot = 0.0;
    for (int rank = 0; rank < s->P; rank++) {
        rank2d *r = &s->r[rank];
        tot += macro_diff(r->dist, s->p.N, r->jStartWoGhosts, r->jEndWoGhosts, r->jStartWoGhosts, r->u, r->v, r->rho);
    }
    return sqrt(tot);
}

/* u, v, rho as the reference's output files hold them: M rows x N columns, rank blocks in order
 * (2D:190-202, 482-495) */
void inplace2d_get_macro(void *h, double *u, double *v, double *rho) {
    inplace2d *s = (inplace2d *)h;
    const int N = s->p.N;
    size_t off = 0;
    for (int rank = 0; rank < s->P; rank++) {
        rank2d *r = &s->r[rank];
        const size_t n = (size_t)N * r->Mj;
        memcpy(u + off, r->u, n * sizeof(double)); memcpy(v + off, r->v, n * sizeof(double));
        memcpy(rho + off, r->rho, n * sizeof(double));
        off += n;
    }
}

/* single-cell entry point for tests/host_device_functions_check.cu */
void oracle2d_collide_cell(const oracle2d_params *p, const double *ft9, double *out9) {
    consts2d c;
    derive(p, &c);
    collide(&c, ft9, out9);
}
