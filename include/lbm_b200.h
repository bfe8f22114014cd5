/*
 * lbm_b200.h -- C ABI of the B200-native lid-driven-cavity Lattice-Boltzmann hot path.
 *
 * Drop-in boundary for the per-timestep collide-and-stream update of
 *   tjburrows/CFD-codes: "3D Lid-driven Cavity with Lattice Boltzmann Method/main.c"  ("3D:")
 * and the D2Q9 variant of the same kernel
 *   "2D Lid-driven Cavity with Lattice Boltzmann Method/main.cpp"                      ("2D:")
 *
 * The reference has no plugin/FFI layer: its hot path is five plain C functions over
 * process-global arrays (3D:50-72).  Each entry point below names the reference function(s) it
 * replaces.  Plain C types only (pointers, sizes, doubles); no CUDA or torch types cross this
 * boundary.  All functions return LBM_OK (0) or an LBM_ERR_* code; lbm_last_error() gives the
 * message of the calling thread's last failure.  There is no CPU fallback: without a CUDA
 * device lbm_create() fails with LBM_ERR_CUDA.
 *
 * Threading: one host thread drives one context (= one GPU = one z-slab).  Multi-GPU runs use
 * one context per GPU, either one process per GPU (torchrun / bench.py) or one thread per GPU
 * (host/main.c); the contexts of a run share an NCCL communicator created from nccl_id.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_OK            0
#define LBM_ERR_INVALID   1 /* bad parameter / bad call order */
#define LBM_ERR_CUDA      2 /* CUDA runtime failure (incl. "no device") */
#define LBM_ERR_UNSTABLE  3 /* a negative equilibrium was computed: the reference prints
                               "Error: negative feq.  Therefore, unstable." and exit(1)s (3D:687-692) */
#define LBM_ERR_NCCL      4
#define LBM_ERR_NOMEM     5

#define LBM_NCCL_ID_BYTES 128
#define LBM_PEER_BLOB_BYTES 320

typedef struct lbm_ctx lbm_ctx;

/* Run-time versions of the reference's compile-time parameters (3D:7-15, 3D:22; 2D:12-17). */
typedef struct lbm_params {
    int    dim;            /* 3 = D3Q19 cube N^3 (3D main.c);  2 = D2Q9 N x M (2D main.cpp)          */
    int    N;              /* 3D:8  grid points per side        / 2D:13 width                        */
    int    M;              /* 2D:14 height (ignored for dim=3)                                       */
    double Re;             /* 3D:7  (an int macro there: pass an integral value) / 2D:12             */
    double UMAX;           /* 3D:9  lid speed                   / 2D:15 Ulid                         */
    int    METHOD;         /* 3D:12 0 = BGK, 1 = TRT            (dim=2: TRT only, 2D:367-413)        */
    double GAMMA;          /* 3D:13 preconditioning coefficient (dim=2: ignored)                     */
    double MAGIC;          /* 3D:22 TRT magic parameter         / 2D:25                              */
    int    DIAG;           /* 3D:14 1 = lid driven along the x-z diagonal (dim=2: ignored)           */
    /* decomposition: rank r of nranks owns planes [r*n/nranks, (r+1)*n/nranks) of the slowest axis
     * (z for dim=3 as in LU4 3D:48; y rows for dim=2 as in 2D:97-125).                               */
    int    rank;
    int    nranks;
    int    device;         /* CUDA device ordinal; -1 = the thread's current device                  */
    int    track_residual; /* 1: keep the previous step's distributions (one extra lattice) so that
                              lbm_residual() is valid after every lbm_step(); 0: lbm_residual fails  */
    void  *stream;         /* optional caller-owned cudaStream_t for the main stream (NULL: own one) */
    unsigned char nccl_id[LBM_NCCL_ID_BYTES]; /* nranks > 1: the same lbm_nccl_unique_id() on all ranks */
} lbm_params;

/* Derived constants (evaluated in the reference's macro order, 3D:19-23,30; 2D:22-26) and layout. */
typedef struct lbm_info {
    double NU, TAU, OMEGA, OMEGAm, OmOMEGA;
    double ulid, vlid, wlid;      /* 3D:183-191 */
    int    Q;                     /* 19 or 9 */
    int    k0, k1;                /* owned plane range of the slowest axis */
    long long cells_local;        /* owned cells */
    long long cells_global;
    long long steps_done;         /* iterations executed since creation / lbm_set_f */
    long long kernel_launches;    /* kernels launched by this context so far */
    long long device_bytes;       /* device memory currently allocated by this context */
    double plain_step_ms;         /* accumulated device time (CUDA events on the launching stream) of the
                                     iterations run by the plain step kernel inside lbm_step (dim=3) ... */
    long long plain_step_iterations; /* ... and how many iterations that covers */
    int pair_enabled;             /* 1: plain iterations run two per pass over HBM (csrc/lbm3d_pair.cuh) */
    int pair_variant;             /* ... with this tile variant (7 = direct pulls, 8/9/10 = TMA-staged) */
} lbm_info;

/* Fills *p with the reference's defaults: dim=3, N=50, Re=100, UMAX=0.1, METHOD=1 (TRT), GAMMA=1,
 * MAGIC=0.25, DIAG=0 (3D:7-15,22); rank 0 of 1, device -1, track_residual=1. */
void lbm_default_params(lbm_params *p);

/* ncclGetUniqueId for a multi-GPU run; call on one rank and distribute the bytes. */
int lbm_nccl_unique_id(unsigned char out[LBM_NCCL_ID_BYTES]);

/* Replaces allocate() + initialize() (3D:135-147, 172-206; 2D:127-158): device allocation,
 * derived constants, f = feq(rho=1, u=0) everywhere, lid velocity.  Collective when nranks > 1. */
int lbm_create(const lbm_params *p, lbm_ctx **out);

/* Optional direct NVLink halo for dim=3 multi-GPU runs (default transport: NCCL send/recv).
 * Every rank calls lbm_peer_export and hands its blob to its two z-neighbours (any transport);
 * every rank then calls lbm_peer_connect with the blobs of rank-1 / rank+1 (NULL at the ends),
 * all at the same iteration count.  From then on the kernel that computes a slab-boundary plane also
 * stores the five face-crossing populations straight into the neighbour GPU's ghost plane (CUDA IPC /
 * peer access) and raises an arrival flag there: compute + exchange in one kernel, no send/recv.
 * In this mode lbm_get_f / lbm_macrovars are collective calls when track_residual is 0. */
int lbm_peer_export(lbm_ctx *c, unsigned char blob[LBM_PEER_BLOB_BYTES]);
int lbm_peer_connect(lbm_ctx *c, const unsigned char *below_blob, const unsigned char *above_blob);

/* Replaces n iterations of `macrocollideandstream(); HechtBC(f2); swap(&f1,&f2);`
 * (3D:94-100, 106-123; 2D: one streamCollide sweep of all rows per iteration, 2D:160-178).
 * Asynchronous on the context's streams.  Returns LBM_ERR_UNSTABLE if an earlier poll
 * (lbm_residual / lbm_sync) found the negative-feq flag set. */
int lbm_step(lbm_ctx *c, long long n);

/* lbm_step + CUDA-event timing on the launching stream (ms, device time, synchronises). */
int lbm_step_timed(lbm_ctx *c, long long n, float *ms);

/* CUDA-event stopwatch on the context's launching stream (what bench.py times with):
 * start synchronises and records; stop records after all queued work (incl. the halo stream)
 * and returns the device time in ms. */
int lbm_timer_start(lbm_ctx *c);
int lbm_timer_stop(lbm_ctx *c, float *ms);

/* Replaces diff(f2, f1, NQ) (3D:96, 112, 678-685): sqrt(sum over all slots of (f(t) - f(t-1))^2)
 * for the most recent iteration; summed over ranks.  Requires track_residual.  Synchronises and
 * polls the negative-feq flag (returns LBM_ERR_UNSTABLE if set; *out is still written). */
int lbm_residual(lbm_ctx *c, double *out);

/* Replaces the reference's convergence loop (3D:103-124): up to max_blocks blocks of `skip`
 * iterations; after each block df = diff(f2, f1, NQ) * df0 is evaluated and the loop stops when
 * df < thresh (3D:119-122), when a negative equilibrium was seen, or after max_blocks.  On one GPU
 * (skip a multiple of 4) the WHOLE loop runs as one CUDA graph with a device-side while condition:
 * the host does not synchronise with the GPU between the blocks; `on_block` is called on the calling
 * thread for every finished block, in order, as its record arrives in mapped host memory (the place
 * for the reference's printf block, 3D:113-117).  Otherwise (several GPUs, other skip, option
 * "converge_on_device" = 0) the same loop runs with one lbm_residual per block.  *blocks_done blocks
 * were executed; *df_last is the last df.  Returns LBM_ERR_UNSTABLE like lbm_residual. */
typedef struct lbm_block_record {
    long long iteration;   /* iteration index t of this block's residual check */
    double df;             /* diff(f2, f1, NQ) * df0 */
    int status;            /* LBM_OK or LBM_ERR_UNSTABLE */
    int last;              /* 1: the loop ended with this block */
} lbm_block_record;
typedef void (*lbm_block_fn)(const lbm_block_record *rec, void *user);
int lbm_converge(lbm_ctx *c, long long skip, double thresh, long long max_blocks, double df0, lbm_block_fn on_block,
                 void *user, long long *blocks_done, double *df_last);

/* Direct access to the reference's `f2` (the distributions after the last iteration, post-BC,
 * pre-collision): writes this rank's planes into a FULL array of Q*cells_global doubles in the
 * reference's LU4 order (3D:48: q*N^3 + k*N^2 + j*N + i); other ranks' planes are untouched. */
int lbm_get_f(lbm_ctx *c, double *host_full);

/* Restart from host distributions in LU4 order (full arrays; this rank reads its planes).
 * f_prev may be NULL; if given it is the reference's other buffer (the one about to be
 * overwritten), from which the never-streamed slot f14 at x = N-1 is taken (3D:421). */
int lbm_set_f(lbm_ctx *c, const double *host_f, const double *host_f_prev);

/* Asynchronous halves of lbm_set_f (dim=3), for hosts that feed a SEQUENCE of states and want the
 * PCIe upload of the next state to overlap the iterations of the current one:
 *   lbm_set_f_async  enqueues the host->device copies on the context's copy stream into a staging
 *                    buffer and returns at once (host arrays: pinned memory, valid until the copy is
 *                    done, i.e. until lbm_io_wait or the lbm_set_f_commit that follows has been
 *                    waited on with lbm_sync);
 *   lbm_set_f_commit makes the staged state the current one (stream-ordered after the upload;
 *                    collective when nranks > 1, like lbm_set_f).
 * lbm_set_f == lbm_set_f_async + lbm_set_f_commit + wait. */
int lbm_set_f_async(lbm_ctx *c, const double *host_f, const double *host_f_prev);
int lbm_set_f_commit(lbm_ctx *c);

/* Replaces macrovars(f2) + calcvmag() (3D:365-366, 695-738): N^3 doubles each in LU3 order
 * (3D:47), this rank's planes only; any pointer may be NULL.  Velocities are divided by rho
 * (not multiplied by a reciprocal), as 3D:713-715. */
int lbm_macrovars(lbm_ctx *c, double *u1, double *u2, double *u3, double *rho, double *vmag);

/* lbm_macrovars without the final wait (dim=3): the fields are computed in stream order and
 * downloaded on the copy stream; the host arrays are complete after lbm_io_wait. */
int lbm_macrovars_async(lbm_ctx *c, double *u1, double *u2, double *u3, double *rho, double *vmag);
/* Waits for the transfers queued by lbm_set_f_async / lbm_macrovars_async. */
int lbm_io_wait(lbm_ctx *c);

/* Waits for all work of the context and polls the negative-feq flag. */
int lbm_sync(lbm_ctx *c);

int lbm_get_info(lbm_ctx *c, lbm_info *out);

/* Options: "block_x" / "block_y" (thread-block shape of the step kernel), "track_residual" (0/1 at
 * run time), "pair" (two iterations per pass over HBM, csrc/lbm3d_pair.cuh: -1 automatic, 0 off, 1 on;
 * on several GPUs switching it on is a collective call), "pair_variant" / "pair_lz" (its tile variant
 * and planes per z segment), "graphs" (CUDA-graph replay of launch-bound grids), "vec2", "halo_timeout_ms"
 * (direct halo: how long a GPU waits for a neighbour before reporting an error; ranks must stay within
 * that skew of each other, default 600 000), "host_local_layout" (1: the host arrays of lbm_get_f / lbm_set_f / lbm_macrovars are
 * compact local slabs [q][k1-k0][N][N] instead of full LU4 / LU3 arrays -- for multi-process runs
 * where no rank holds the whole field). */
int lbm_set_option(lbm_ctx *c, const char *key, long long value);

/* Replaces freemem() (3D:150-160). */
int lbm_destroy(lbm_ctx *c);

const char *lbm_last_error(void);

/* Library/build identification, e.g. "lbm_b200 sm_100a fmad=false". */
const char *lbm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
