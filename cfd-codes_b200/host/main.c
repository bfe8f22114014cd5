/*
 * host/main.c -- C host program of the B200 cavity solver: the drop-in for the reference's
 *   "3D Lid-driven Cavity with Lattice Boltzmann Method/main.c"  ("3D:" below).
 *
 * Same parameter surface (now run-time flags with the reference's names and defaults, 3D:7-15, 18, 22,
 * 43), same stdout (banner 3D:80-88, residual blocks 3D:97-98, 113-117), same convergence logic
 * (3D:93-124) and the same Tecplot output file (3D:174, 363-382).  All lattice work goes through
 * the C ABI of include/lbm_b200.h; this file contains no numerics except printf formatting.
 *
 *   lbm3d_cavity [--N 50] [--Re 100] [--UMAX 0.1] [--SAVETXT 1] [--THRESH 1E-12] [--METHOD 1]
 *                [--GAMMA 1.0] [--DIAG 0] [--MAGIC 0.25] [--MAXITER 1E8] [--skip 500] [--gpus 1] [--halo 0]
 *
 * --gpus P (the reference's THREADS knob, 3D:15) runs P z-slabs, one host thread + one context
 * per GPU, sharing an NCCL communicator.  --halo 1 switches the slab exchange from NCCL send/recv to
 * the fused direct-NVLink kernel (lbm_peer_export / lbm_peer_connect).
 *
 * Deviations from the reference's stdout, all deliberate:
 *   - "Threads used:" prints the GPU count;
 *   - LUPS is computed in double (the reference's `N3 * skip` overflows int for N >= 163, 3D:117);
 *   - a negative equilibrium is detected at the next residual check (<= skip iterations late)
 *     instead of immediately; message and exit status are the reference's (3D:687-692).
 * Kept on purpose: when the loop ends by reaching MAXITER without converging, the reference has
 * already swapped its buffers, so datwrite's macrovars(f2) reads the state of iteration MAXITER-2
 * (3D:104-127); this program writes that same state (the fields are taken one iteration before the
 * last).  With MAXITER <= 1 the reference would print the initial state; here the state after
 * iteration 0 is written.
 *
 * Errors: a failing library call makes the worker print the reference's message and return its
 * status to main(), which exits 1.  With --gpus > 1 the other workers share an NCCL communicator
 * with the failed rank and could wait for it forever, so the process is ended with _exit(1) (no
 * atexit handlers or destructors run while sibling threads are still inside CUDA/NCCL).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "fmt10.h"
#include "lbm_b200.h"

typedef struct {
    int N, Re, SAVETXT, METHOD, DIAG, skip, gpus, halo;
    double UMAX, THRESH, GAMMA, MAGIC, MAXITER;
} options;

typedef struct {
    int rank;
    const options *opt;
    const unsigned char *nccl_id;
    pthread_barrier_t *bar;
    double *u1, *u2, *u3, *vmag; /* shared N^3 arrays (SAVETXT) */
    unsigned char (*blobs)[LBM_PEER_BLOB_BYTES]; /* --halo 1: one export blob per rank */
    int status;
} worker;

typedef struct {
    int root;
    long long skip;
    double cells, start;
} block_printer;

static double now(void);

/* The reference's printf block after every residual check, 3D:113-117.  Called by lbm_converge on the
 * worker's own thread as each block finishes on the GPU. */
static void print_block(const lbm_block_record *r, void *user) {
    block_printer *bp = (block_printer *)user;
    const double stop = now() - bp->start;
    if (bp->root) {
        printf("\nIteration %lld:\n", r->iteration);
        printf("df/df0:\t%.3e\n", r->df);
        printf("Time:\t%.3e s\n", stop);
        printf("LUPS:\t%.3e s\n", bp->cells * (double)bp->skip / stop);
        fflush(stdout);
    }
    bp->start = now();
}

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int g_gpus = 1;

/* Report a failed library call the way the reference reports its own errors (3D:687-692, 335-344).
 * Single GPU: returns so that the worker can hand the status to main().  Several GPUs: see "Errors"
 * in the header comment. */
static void report_lbm(int rc) {
    if (rc == LBM_ERR_UNSTABLE) {
        printf("Error: negative feq.  Therefore, unstable.\n"); /* 3D:689 */
    } else {
        printf("Error: %s\n", lbm_last_error());
    }
    fflush(stdout);
    if (g_gpus > 1) _exit(1);
}
#define LBM_TRY(call)                                   \
    do {                                                \
        if ((rc = (call))) {                            \
            report_lbm(rc);                             \
            if (ctx) lbm_destroy(ctx);                  \
            w->status = 1;                              \
            return NULL;                                \
        }                                               \
    } while (0)

/* Writes the Tecplot file, 3D:363-382: i outer, j, k inner; x = y = z = linspace(0, N-1, N) (3D:175).
 * Same bytes as the reference's fprintf("%.10f, ...") loop, produced by an exact integer formatter
 * (fmt10.h) on all host threads: one i-slice (N^2 lines) per thread, written in order. */
static void datwrite(const options *o, const double *u1, const double *u2, const double *u3, const double *vmag) {
    char filename[160];
    snprintf(filename, sizeof filename, "Solution_n=%dRe=%d_DIAG=%d_%dRT_M=%.3f.dat", o->N, o->Re, o->DIAG,
             o->METHOD + 1, o->MAGIC); /* 3D:174 */
    FILE *fp = fopen(filename, "w");
    if (fp == NULL) {
        printf("Error opening file!\n");
        exit(1);
    }
    const int N = o->N;
    const size_t N1 = (size_t)N, N2 = N1 * N1;
    fprintf(fp, "TITLE=\"%s\" VARIABLES=\"x\", \"y\", \"z\", \"%s\", \"%s\", \"%s\", \"vmag\" ZONE T=\"%s\" I=%d J=%d K=%d F=POINT\n",
            filename, "u", "v", "w", filename, N, N, N);
    int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    if (T > N) T = N;
    const size_t line_max = 7 * 24 + 16;
    char **bufs = malloc((size_t)T * sizeof(char *));
    size_t *lens = calloc((size_t)T, sizeof(size_t));
    for (int t = 0; t < T; t++) {
        bufs[t] = malloc(N2 * line_max);
        if (!bufs[t]) { printf("Error allocating memory!\n"); exit(1); }
    }
    for (int i0 = 0; i0 < N; i0 += T) {
        const int i1 = i0 + T < N ? i0 + T : N;
#pragma omp parallel for schedule(static, 1)
        for (int i = i0; i < i1; i++) {
            char *p = bufs[i - i0];
            for (int j = 0; j < N; j++)
                for (int k = 0; k < N; k++) {
                    const size_t c = (size_t)k * N2 + (size_t)j * N1 + (size_t)i; /* LU3, 3D:47 */
                    /* linspace(0, N-1, N): start + i*(stop-start)/(num-1) is exactly i (3D:328) */
                    const double v[7] = {(double)i, (double)j, (double)k, u1[c], u2[c], u3[c], vmag[c]};
                    for (int a = 0; a < 7; a++) {
                        p += fmt10(p, v[a]);
                        if (a < 6) { *p++ = ','; *p++ = ' '; }
                    }
                    *p++ = '\n';
                }
            lens[i - i0] = (size_t)(p - bufs[i - i0]);
        }
        for (int i = i0; i < i1; i++)
            if (fwrite(bufs[i - i0], 1, lens[i - i0], fp) != lens[i - i0]) { printf("Error writing file!\n"); exit(1); }
    }
    for (int t = 0; t < T; t++) free(bufs[t]);
    free(bufs); free(lens);
    fclose(fp);
}

static void *run(void *arg) {
    worker *w = (worker *)arg;
    const options *o = w->opt;
    const int root = (w->rank == 0);
    lbm_params p;
    lbm_default_params(&p);
    p.dim = 3; p.N = o->N; p.Re = o->Re; p.UMAX = o->UMAX; p.METHOD = o->METHOD;
    p.GAMMA = o->GAMMA; p.MAGIC = o->MAGIC; p.DIAG = o->DIAG;
    p.rank = w->rank; p.nranks = o->gpus; p.device = w->rank; p.track_residual = 1;
    if (o->gpus > 1) memcpy(p.nccl_id, w->nccl_id, LBM_NCCL_ID_BYTES);

    lbm_ctx *ctx = NULL;
    int rc;
    LBM_TRY(lbm_create(&p, &ctx)); /* allocate(); initialize();  3D:76-78 */
    lbm_info info;
    lbm_get_info(ctx, &info);
    if (o->halo == 1 && o->gpus > 1) { /* fused direct-NVLink halo instead of NCCL send/recv */
        LBM_TRY(lbm_peer_export(ctx, w->blobs[w->rank]));
        pthread_barrier_wait(w->bar);
        LBM_TRY(lbm_peer_connect(ctx, w->rank > 0 ? w->blobs[w->rank - 1] : NULL,
                                 w->rank + 1 < o->gpus ? w->blobs[w->rank + 1] : NULL));
        pthread_barrier_wait(w->bar);
    }

    if (root) { /* 3D:80-88 */
        printf("Re =\t%d\n", o->Re);
        printf("Re_g =\t%.3e\n", 3.0 * o->UMAX / (info.TAU - 0.5));
        printf("U =\t%.3e\n", o->UMAX);
        printf("N =\t%d\n", o->N);
        printf("tau =\t%.3e\n", info.TAU);
        printf("nu =\t%.3e\n", info.NU);
        printf("M =\t%.3e\n", o->UMAX * sqrt(3));
        if (o->GAMMA != 1.0) printf("M* =\t%.3e\n", o->UMAX * sqrt(3) / sqrt(o->GAMMA));
    }

    /* First iteration, 3D:94-100 */
    double df = 0.0, df0;
    LBM_TRY(lbm_step(ctx, 1));
    LBM_TRY(lbm_residual(ctx, &df0));
    if (root) {
        printf("\nIteration 0:\n");
        printf("df:\t%.3e\n", df0);
    }
    df0 = 1.0 / df0;

    /* Rest of iterations, 3D:103-124: iteration t is checked when t % skip == 0.  All blocks that end
     * strictly before the last iteration run inside lbm_converge -- on one GPU a single device-side
     * while loop, no host synchronisation between the blocks.  The block that contains the last
     * iteration (MAXITER exit, see header) is stepped by hand.  `fields_taken`: the output fields were
     * already read one iteration before the last. */
    const double cells = (double)o->N * (double)o->N * (double)o->N;
    block_printer bp = {root, o->skip, cells, now()};
    long long t = 1;
    const long long last = (long long)o->MAXITER - 1;
    int fields_taken = 0, converged = 0;
    {
        const long long whole = last >= 1 ? (last - 1) / o->skip : 0; /* checks at skip, 2*skip, ... < last */
        long long done = 0;
        LBM_TRY(lbm_converge(ctx, o->skip, o->THRESH, whole, df0, print_block, &bp, &done, &df));
        t += done * o->skip;
        converged = (done > 0 && df < o->THRESH);
    }
    while (!converged && t <= last) {
        long long T = ((t + o->skip - 1) / o->skip) * o->skip; /* next check iteration */
        const long long stop_at = T >= last ? last : T;
        /* run up to iteration last-1, take the fields the reference would print if this block
         * does not converge, then do the final iteration */
        if (stop_at - t > 0) LBM_TRY(lbm_step(ctx, stop_at - t));
        if (o->SAVETXT == 1) {
            LBM_TRY(lbm_macrovars(ctx, w->u1, w->u2, w->u3, NULL, w->vmag));
            fields_taken = 1;
        }
        LBM_TRY(lbm_step(ctx, 1));
        if (T > last) { /* no residual check at the final iteration */
            LBM_TRY(lbm_sync(ctx));
            break;
        }
        LBM_TRY(lbm_residual(ctx, &df));
        df *= df0;
        lbm_block_record r = {T, df, 0, 1};
        print_block(&r, &bp);
        if (df < o->THRESH) fields_taken = 0; /* converged: the newest state is printed (3D:119-122) */
        break;
    }

    /* Output to text file, 3D:127 */
    if (o->SAVETXT == 1) {
        if (!fields_taken) LBM_TRY(lbm_macrovars(ctx, w->u1, w->u2, w->u3, NULL, w->vmag));
        if (o->gpus > 1) pthread_barrier_wait(w->bar);
        if (root) datwrite(o, w->u1, w->u2, w->u3, w->vmag);
    }
    lbm_destroy(ctx); /* freemem(), 3D:130 */
    w->status = 0;
    return NULL;
}

static int flag(int argc, char **argv, int *i, const char *name, const char **val) {
    if (strcmp(argv[*i], name) != 0) return 0;
    if (*i + 1 >= argc) {
        fprintf(stderr, "%s needs a value\n", name);
        exit(2);
    }
    *val = argv[++*i];
    return 1;
}

int main(int argc, char **argv) {
    /* defaults = the reference's #defines, 3D:7-15, 18, 22, 43 */
    options o = {.N = 50, .Re = 100, .SAVETXT = 1, .METHOD = 1, .DIAG = 0, .skip = 500, .gpus = 1, .halo = 0,
                 .UMAX = 0.1, .THRESH = 1E-12, .GAMMA = 1.0, .MAGIC = 0.25, .MAXITER = 1E8};
    for (int i = 1; i < argc; i++) {
        const char *v;
        if (flag(argc, argv, &i, "--N", &v)) o.N = atoi(v);
        else if (flag(argc, argv, &i, "--Re", &v)) o.Re = atoi(v);
        else if (flag(argc, argv, &i, "--UMAX", &v)) o.UMAX = atof(v);
        else if (flag(argc, argv, &i, "--SAVETXT", &v)) o.SAVETXT = atoi(v);
        else if (flag(argc, argv, &i, "--THRESH", &v)) o.THRESH = atof(v);
        else if (flag(argc, argv, &i, "--METHOD", &v)) o.METHOD = atoi(v);
        else if (flag(argc, argv, &i, "--GAMMA", &v)) o.GAMMA = atof(v);
        else if (flag(argc, argv, &i, "--DIAG", &v)) o.DIAG = atoi(v);
        else if (flag(argc, argv, &i, "--MAGIC", &v)) o.MAGIC = atof(v);
        else if (flag(argc, argv, &i, "--MAXITER", &v)) o.MAXITER = atof(v);
        else if (flag(argc, argv, &i, "--skip", &v)) o.skip = atoi(v);
        else if (flag(argc, argv, &i, "--gpus", &v) || flag(argc, argv, &i, "--THREADS", &v)) o.gpus = atoi(v);
        else if (flag(argc, argv, &i, "--halo", &v)) o.halo = atoi(v);
        else {
            fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    if (o.gpus < 1 || o.skip < 1) {
        fprintf(stderr, "bad --gpus / --skip\n");
        return 2;
    }
    printf("Threads used:\t%d\n", o.gpus); /* 3D:210 (here: GPUs) */
    g_gpus = o.gpus;

    unsigned char id[LBM_NCCL_ID_BYTES];
    memset(id, 0, sizeof id);
    if (o.gpus > 1 && lbm_nccl_unique_id(id)) {
        printf("Error: %s\n", lbm_last_error());
        return 1;
    }

    double *u1 = NULL, *u2 = NULL, *u3 = NULL, *vmag = NULL;
    if (o.SAVETXT == 1) {
        const size_t n3 = (size_t)o.N * o.N * o.N;
        u1 = malloc(n3 * sizeof(double)); u2 = malloc(n3 * sizeof(double));
        u3 = malloc(n3 * sizeof(double)); vmag = malloc(n3 * sizeof(double));
        if (!u1 || !u2 || !u3 || !vmag) {
            printf("Error allocating memory!\n"); /* 3D:339 */
            return 1;
        }
    }
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)o.gpus);
    unsigned char(*blobs)[LBM_PEER_BLOB_BYTES] = calloc((size_t)o.gpus, LBM_PEER_BLOB_BYTES);
    worker *ws = calloc((size_t)o.gpus, sizeof(worker));
    pthread_t *th = calloc((size_t)o.gpus, sizeof(pthread_t));
    for (int r = 0; r < o.gpus; r++) {
        ws[r] = (worker){.rank = r, .opt = &o, .nccl_id = id, .bar = &bar, .u1 = u1, .u2 = u2, .u3 = u3,
                         .vmag = vmag, .blobs = blobs, .status = 1};
        if (r > 0) pthread_create(&th[r], NULL, run, &ws[r]);
    }
    run(&ws[0]);
    for (int r = 1; r < o.gpus; r++) pthread_join(th[r], NULL);
    int status = 0;
    for (int r = 0; r < o.gpus; r++) status |= ws[r].status;
    free(u1); free(u2); free(u3); free(vmag); free(ws); free(th); free(blobs);
    return status ? 1 : 0;
}
