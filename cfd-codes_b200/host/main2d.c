/*
 * host/main2d.c -- C host program of the D2Q9 variant: the counterpart of the reference's
 *   "2D Lid-driven Cavity with Lattice Boltzmann Method/main.cpp"  ("2D:" below)
 * on B200s.  Same parameter surface as run-time flags (2D:12-17, 21, 25, 31), same stdout
 * (banner 2D:74-81, rank lines 2D:126-132, table 2D:464, 470-472), same convergence logic
 * (macroVars + convergence after iterations 1, prntInt+1, ..., 2D:160-187, 452-478) and the same
 * output files Re=%.0f_N=%d_M=%d_{x,y,u,v,rho}.bin (raw doubles, M rows x N columns; 2D:190-202,
 * 482-495; readable by the reference's plot.py).
 *
 *   lbm2d_cavity [--Re 100] [--N 100] [--M 300] [--Ulid 0.1] [--SAVE 1] [--THRESH 2E-9]
 *                [--MAXITER 1E6] [--prntInt 1000] [--MAGIC 0.25] [--gpus 1]
 *
 * Deliberate differences: the update is the time-consistent two-lattice version of the reference's
 * in-place sweep (same steady state) so iteration counts differ (94 000 vs 44 000 at
 * the defaults); any number of GPUs >= 1 (the reference insists on >= 3 MPI ranks, 2D:69-72);
 * MLUPS is computed in double (the reference's `prntInt * N * M` overflows int, 2D:472).
 * Kept on purpose: the files hold u, v, rho of the LAST CHECK (2D:180-198), also when the loop ends
 * at MAXITER some iterations after it.  A failing library call prints the message and makes main()
 * return 1 (several GPUs: _exit(1), the sibling workers share an NCCL communicator with the failed one).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "lbm_b200.h"

typedef struct {
    int N, M, SAVE, prntInt, gpus;
    double Re, Ulid, THRESH, MAXITER, MAGIC;
} options;

typedef struct {
    int rank;
    const options *opt;
    const unsigned char *nccl_id;
    pthread_barrier_t *bar;
    double *u, *v, *rho;
    int status;
} worker;

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int g_gpus = 1;

static void report_lbm(void) {
    printf("Error: %s\n", lbm_last_error());
    fflush(stdout);
    if (g_gpus > 1) _exit(1);
}
#define LBM_TRY(call)                      \
    do {                                   \
        if ((call)) {                      \
            report_lbm();                  \
            if (ctx) lbm_destroy(ctx);     \
            w->status = 1;                 \
            return NULL;                   \
        }                                  \
    } while (0)

/* outputScalar, 2D:482-495 (single writer: the slabs were gathered into one host array) */
static void output_scalar(const options *o, const char *name, const double *a) {
    char fn[128];
    snprintf(fn, sizeof fn, "Re=%.0f_N=%d_M=%d_%s.bin", o->Re, o->N, o->M, name);
    FILE *fp = fopen(fn, "wb");
    if (!fp) { printf("Error opening file!\n"); exit(1); }
    fwrite(a, sizeof(double), (size_t)o->N * o->M, fp);
    fclose(fp);
}

static void *run(void *arg) {
    worker *w = (worker *)arg;
    const options *o = w->opt;
    const int root = (w->rank == 0);
    lbm_params p;
    lbm_default_params(&p);
    p.dim = 2; p.N = o->N; p.M = o->M; p.Re = o->Re; p.UMAX = o->Ulid; p.METHOD = 1; p.MAGIC = o->MAGIC;
    p.rank = w->rank; p.nranks = o->gpus; p.device = w->rank; p.track_residual = 0;
    if (o->gpus > 1) memcpy(p.nccl_id, w->nccl_id, LBM_NCCL_ID_BYTES);
    lbm_ctx *ctx = NULL;
    LBM_TRY(lbm_create(&p, &ctx));
    lbm_info info;
    lbm_get_info(ctx, &info);
    if (root) { /* 2D:74-81 */
        printf("MPI LBM Simulation of Lid Driven Cavity\n");
        printf("Domain Size: %d x %d\n", o->N, o->M);
        printf("Re\t= %g\n", o->Re);
        printf("nu\t= %g\n", info.NU);
        printf("tau\t= %g\n", info.TAU);
        printf("Umax\t= %g\n", o->Ulid);
    }
    for (int r = 0; r < o->gpus; r++) { /* 2D:126-132 */
        if (r == w->rank)
            printf("Rank %d: %d nodes from y = %3d to y = %3d (%s)\n", w->rank, info.k1 - info.k0, info.k0, info.k1 - 1,
                   w->rank == 0 ? "BOTTOM" : (w->rank == o->gpus - 1 ? "TOP" : "INTERNAL"));
        fflush(stdout);
        if (o->gpus > 1) pthread_barrier_wait(w->bar);
    }

    double diff, df0 = 1.0, df = 1.0;
    double start = now();
    const double cells = (double)o->N * (double)o->M;
    /* 2D:160-187: iteration t, then the check when t % prntInt == 0 */
    long long t = 0;
    int fields_taken = 0;
    const long long maxiter = (long long)o->MAXITER;
    while (t < maxiter) {
        const long long T = (t == 0) ? 0 : ((t + o->prntInt - 1) / o->prntInt) * o->prntInt; /* next check iteration */
        if (T >= maxiter) {
            /* MAXITER ends the run between two checks: the files hold the fields of the last check */
            if (o->SAVE && t > 0) {
                LBM_TRY(lbm_macrovars(ctx, w->u, w->v, NULL, w->rho, NULL));
                fields_taken = 1;
            }
            LBM_TRY(lbm_step(ctx, maxiter - t));
            break;
        }
        LBM_TRY(lbm_step(ctx, T - t + 1));
        LBM_TRY(lbm_residual(ctx, &diff)); /* macroVars + Allreduce + sqrt, 2D:419-460 */
        if (T == 0) {
            df0 = 1.0 / diff;
            if (root) printf("Iteration\tdf/df0\t\tMLUPS\n");
        }
        df = diff * df0;
        const double stop = now();
        if (root) {
            printf("%.3e\t%.3e\t%.1f\n", (double)T, df, (double)o->prntInt * cells / (1E6 * (stop - start)));
            fflush(stdout);
            if (isinf(df)) { fflush(stdout); _exit(1); } /* 2D:473 */
        }
        start = stop;
        if (df < o->THRESH) break;
        t = T + 1;
    }
    if (o->SAVE) {
        /* the reference writes u, v, rho of the LAST check (2D:190-202) = the state just examined */
        if (!fields_taken) LBM_TRY(lbm_macrovars(ctx, w->u, w->v, NULL, w->rho, NULL));
        if (o->gpus > 1) pthread_barrier_wait(w->bar);
        if (root) {
            double *x = malloc((size_t)o->N * o->M * sizeof(double)), *y = malloc((size_t)o->N * o->M * sizeof(double));
            for (int j = 0; j < o->M; j++)
                for (int i = 0; i < o->N; i++) { /* 2D:146-152 */
                    x[(size_t)j * o->N + i] = 0.5 + i;
                    y[(size_t)j * o->N + i] = 0.5 + j;
                }
            output_scalar(o, "x", x); output_scalar(o, "y", y);
            output_scalar(o, "u", w->u); output_scalar(o, "v", w->v); output_scalar(o, "rho", w->rho);
            free(x); free(y);
        }
    }
    lbm_destroy(ctx);
    w->status = 0;
    return NULL;
}

static int flag(int argc, char **argv, int *i, const char *name, const char **val) {
    if (strcmp(argv[*i], name) != 0) return 0;
    if (*i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", name); exit(2); }
    *val = argv[++*i];
    return 1;
}

int main(int argc, char **argv) {
    options o = {.N = 100, .M = 300, .SAVE = 1, .prntInt = 1000, .gpus = 1, .Re = 100, .Ulid = 0.1, .THRESH = 2E-9,
                 .MAXITER = 1E6, .MAGIC = 0.25}; /* 2D:12-17, 21, 25, 31 */
    for (int i = 1; i < argc; i++) {
        const char *v;
        if (flag(argc, argv, &i, "--N", &v)) o.N = atoi(v);
        else if (flag(argc, argv, &i, "--M", &v)) o.M = atoi(v);
        else if (flag(argc, argv, &i, "--Re", &v)) o.Re = atof(v);
        else if (flag(argc, argv, &i, "--Ulid", &v)) o.Ulid = atof(v);
        else if (flag(argc, argv, &i, "--SAVE", &v)) o.SAVE = atoi(v);
        else if (flag(argc, argv, &i, "--THRESH", &v)) o.THRESH = atof(v);
        else if (flag(argc, argv, &i, "--MAXITER", &v)) o.MAXITER = atof(v);
        else if (flag(argc, argv, &i, "--prntInt", &v)) o.prntInt = atoi(v);
        else if (flag(argc, argv, &i, "--MAGIC", &v)) o.MAGIC = atof(v);
        else if (flag(argc, argv, &i, "--gpus", &v)) o.gpus = atoi(v);
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (o.gpus < 1 || o.prntInt < 1) { fprintf(stderr, "bad --gpus / --prntInt\n"); return 2; }
    unsigned char id[LBM_NCCL_ID_BYTES];
    memset(id, 0, sizeof id);
    g_gpus = o.gpus;
    if (o.gpus > 1 && lbm_nccl_unique_id(id)) { printf("Error: %s\n", lbm_last_error()); return 1; }
    const size_t n = (size_t)o.N * o.M;
    double *u = malloc(n * sizeof(double)), *v = malloc(n * sizeof(double)), *rho = malloc(n * sizeof(double));
    if (!u || !v || !rho) { printf("Error allocating memory!\n"); return 1; }
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)o.gpus);
    worker *ws = calloc((size_t)o.gpus, sizeof(worker));
    pthread_t *th = calloc((size_t)o.gpus, sizeof(pthread_t));
    for (int r = 0; r < o.gpus; r++) {
        ws[r] = (worker){.rank = r, .opt = &o, .nccl_id = id, .bar = &bar, .u = u, .v = v, .rho = rho, .status = 1};
        if (r > 0) pthread_create(&th[r], NULL, run, &ws[r]);
    }
    run(&ws[0]);
    for (int r = 1; r < o.gpus; r++) pthread_join(th[r], NULL);
    int status = 0;
    for (int r = 0; r < o.gpus; r++) status |= ws[r].status;
    free(u); free(v); free(rho); free(ws); free(th);
    return status ? 1 : 0;
}
