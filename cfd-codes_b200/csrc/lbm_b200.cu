// lbm_b200.cu -- host side of the C ABI declared in include/lbm_b200.h (contexts, streams, slabs,
// NCCL halo exchange) for the D3Q19 cavity path.  See DESIGN.md for the state machine.
//
// Reference being replaced: /root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c
// ("3D:"): allocate/initialize 3D:135-206, the time loop 3D:94-124, diff 3D:678, macrovars 3D:695.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm3d_kernels.cuh"
#include "lbm2d_kernels.cuh"

// ---- error plumbing ------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(e_ == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                 \
    } while (0)

// ---- NCCL through dlopen (only touched when nranks > 1) ---------------------------------------
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.h) return LBM_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    void *h = nullptr;
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(LBM_ERR_NCCL, "dlopen(libnccl.so.2) failed: %s", dlerror());
#define SYM(field, name)                                                         \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                  \
    if (!g_nccl.field) return fail(LBM_ERR_NCCL, "dlsym(%s) failed", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.h = h;
    return LBM_OK;
}

#define NK(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess)                                                                        \
            return fail(LBM_ERR_NCCL, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// populations that cross a z face (velocities 3D:36-38): cz = +1 and cz = -1
static const int Q_UP[5] = {5, 9, 11, 16, 18};
static const int Q_DN[5] = {6, 10, 12, 15, 17};

// ---- context -----------------------------------------------------------------------------------
struct lbm_ctx {
    lbm_params prm;
    int dev = 0;
    cudaStream_t s_main = nullptr, s_halo = nullptr;
    bool own_main = false;
    cudaEvent_t ev_bnd = nullptr, ev_halo = nullptr, ev_main = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    bool halo_pending = false;  // ev_halo has been recorded and not yet waited on by s_main

    int N = 0, nzl = 0, k0 = 0, k1 = 0;
    long long N2 = 0, PS = 0, FS = 0;
    double NU, TAU, OMEGA, OMEGAm, OmOMEGA, gammainv, ulid, vlid, wlid;

    double *lat_alloc[2] = {nullptr, nullptr};
    long long pad = 0;                    // elements of padding before/after each lattice
    double *lat[2] = {nullptr, nullptr};  // post-collision lattices (= lat_alloc + pad), ghost plane first
    int cur = 0;                          // lat[cur] holds S of the latest iteration
    double *Fbuf = nullptr;               // the reference's f2 of the latest iteration when f_valid
    bool f_valid = false;
    double *side = nullptr;               // [3][nzl][N] stale-f14 history of the x = N-1 face
    double *partials = nullptr;
    long long npartials = 0;
    double *d_red = nullptr;  // [0] local sum, [1] all-reduced sum
    int *d_flag = nullptr;
    double *d_out[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // macrovars staging
    double *h_pin = nullptr;                                         // pinned: [0..1] sums, [2] flag (as double slot)
    bool residual_ready = false;
    bool unstable = false;

    // dim = 2 (D2Q9): N x M, row slabs [j0, j1); lat[] = [9][nyl+2][N]; macroVars state u2/v2/rho2
    int M2 = 0, nyl = 0, j0 = 0, j1 = 0;
    double *u2 = nullptr, *v2 = nullptr, *rho2 = nullptr;
    double *out2[3] = {nullptr, nullptr, nullptr};
    double lid6 = 0.0;

    long long steps = 0, launches = 0, bytes = 0;
    // device time of the plain step iterations (the dominant kernel), measured with CUDA events
    // around each batch of consecutive MODE_STEP iterations inside lbm_step; resolved lazily
    struct Span { cudaEvent_t a, b; long long iters; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
    double plain_ms = 0.0;
    long long plain_iters = 0;
    int block_x = 0, block_y = 0;
    int vec2 = 0;  // 1: plain iterations use the two-cells-per-thread 128-bit kernel (N even)
    // CUDA graphs for launch-bound grids (single GPU): GRAPH_LEN consecutive plain iterations per
    // graph launch.  The kernel arguments of iteration t depend on (lattice parity, t mod 3), a
    // 6-periodic pattern; GRAPH_LEN is a multiple of 6, so a graph captured in one phase leaves the
    // context in the same phase and can be replayed back to back.
    int graphs = -1;                         // -1 auto (small grids), 0 off, 1 on
    cudaGraphExec_t graph_exec[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // direct NVLink halo (lbm_peer_connect): neighbours' lattices / arrival flags mapped here
    bool p2p = false;
    long long p2p_base = 0;               // iteration count at which the direct halo was switched on
    long long *sig = nullptr;             // my arrival flags: [0] written by rank-1, [1] by rank+1
    unsigned int *done_count = nullptr;
    double *peer_lat[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [side 0 = below, 1 = above][lattice]
    long long *peer_sig[2] = {nullptr, nullptr};
    void *peer_ipc[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // opened IPC mappings to close
    int peer_nzl[2] = {0, 0};
    long long peer_PS[2] = {0, 0};
    bool host_local = false;  // host arrays are compact local slabs instead of full LU4/LU3 arrays
    ncclComm_t comm = nullptr;
};

template <typename T>
static int dev_alloc(lbm_ctx *c, T **p, long long count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (size_t)count * sizeof(T));
    if (e != cudaSuccess)
        return fail(LBM_ERR_NOMEM, "cudaMalloc(%lld bytes): %s", count * (long long)sizeof(T), cudaGetErrorString(e));
    c->bytes += count * (long long)sizeof(T);
    *p = (T *)q;
    return LBM_OK;
}

static void choose_block(const lbm_ctx *c, dim3 *blk, dim3 *grd_xy) {
    int bx = c->block_x, by = c->block_y;
    // 128-thread blocks measured best on B200 (profiles/): 4-5 blocks per SM at <= 128 registers.
    if (bx <= 0) {
        bx = ((c->N + 31) / 32) * 32;
        if (bx > 128) bx = 128;
        // prefer a divisor of the row so that no lanes idle (N = 512 -> 128, N = 320 -> 64)
        if (c->N > 128) {
            for (int cand = 128; cand >= 64; cand -= 32)
                if (c->N % cand == 0) { bx = cand; break; }
        }
    }
    if (by <= 0) {
        by = 128 / bx;
        if (by < 1) by = 1;
        if (by > c->N) by = c->N;
    }
    *blk = dim3(bx, by, 1);
    *grd_xy = dim3((c->N + bx - 1) / bx, (c->N + by - 1) / by, 1);
}

// velocities 3D:36-38
static const int CX19[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0};
static const int CY19[19] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1};
static const int CZ19[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1};

static void kp_set_lattices(const lbm_ctx *c, lbm3d::KP *p, const double *src, double *dst) {
    for (int q = 0; q < lbm3d::Q; q++) {
        const long long shift = (long long)CX19[q] + (long long)CY19[q] * c->N + (long long)CZ19[q] * c->N2;
        p->srcq[q] = src ? src + (long long)q * c->PS - shift : nullptr;
        p->dstq[q] = dst ? dst + (long long)q * c->PS : nullptr;
    }
}

static lbm3d::KP make_kp(const lbm_ctx *c) {
    lbm3d::KP p;
    memset(&p, 0, sizeof(p));
    p.N = c->N; p.M = c->N - 1; p.Mz = c->N - 1;
    p.nzl = c->nzl; p.k0 = c->k0;
    p.kl0 = 0; p.kstride = 1; p.partial_base = 0;
    p.N2 = c->N2; p.PS = c->PS; p.FS = c->FS;
    p.omega = c->OMEGA; p.omegam = c->OMEGAm; p.omomega = c->OmOMEGA; p.gammainv = c->gammainv;
    p.ulid = c->ulid; p.vlid = c->vlid; p.wlid = c->wlid;
    p.vfrac = 1.0 / (c->vlid + 1.0);  // 3D:386
    p.third = 1.0 / 3.0;              // 3D:42
    p.sixth = 1.0 / 6.0;              // 3D:41
    p.flag = c->d_flag;
    p.partials = c->partials;
    p.Fbuf = c->Fbuf;
    return p;
}

struct PeerBlob {   // what lbm_peer_export hands to the neighbours (<= LBM_PEER_BLOB_BYTES)
    int magic, pid, dev, nzl;
    long long PS, pad;
    cudaIpcMemHandle_t lat[2], sig;
    void *raw_lat[2], *raw_sig;
};
static_assert(sizeof(PeerBlob) <= LBM_PEER_BLOB_BYTES, "PeerBlob too large");

template <int MODE>
static void launch_step_peer(lbm_ctx *c, const lbm3d::KP &p, dim3 grid, dim3 blk, cudaStream_t st) {
    const bool g1 = (c->prm.GAMMA == 1.0);
    if (c->prm.METHOD == 1) {
        if (g1) lbm3d::k_step<1, MODE, true, true><<<grid, blk, 0, st>>>(p);
        else    lbm3d::k_step<1, MODE, false, true><<<grid, blk, 0, st>>>(p);
    } else {
        if (g1) lbm3d::k_step<0, MODE, true, true><<<grid, blk, 0, st>>>(p);
        else    lbm3d::k_step<0, MODE, false, true><<<grid, blk, 0, st>>>(p);
    }
    c->launches++;
}

template <int MODE>
static void launch_step_mode(lbm_ctx *c, const lbm3d::KP &p, dim3 grid, dim3 blk, cudaStream_t st) {
    const bool g1 = (c->prm.GAMMA == 1.0);  // x * (1.0 / 1.0) == x: the specialised kernel drops the multiply
    if (c->prm.METHOD == 1) {
        if (g1) lbm3d::k_step<1, MODE, true><<<grid, blk, 0, st>>>(p);
        else    lbm3d::k_step<1, MODE, false><<<grid, blk, 0, st>>>(p);
    } else {
        if (g1) lbm3d::k_step<0, MODE, true><<<grid, blk, 0, st>>>(p);
        else    lbm3d::k_step<0, MODE, false><<<grid, blk, 0, st>>>(p);
    }
    c->launches++;
}

static void launch_step(lbm_ctx *c, int mode, const lbm3d::KP &p, dim3 grid, dim3 blk, cudaStream_t st) {
    if (mode == lbm3d::MODE_STEP && c->vec2 && (c->N % 2) == 0) {
        const bool g1 = (c->prm.GAMMA == 1.0);
        dim3 g2((c->N / 2 + blk.x - 1) / blk.x, grid.y, grid.z);
        if (c->prm.METHOD == 1) {
            if (g1) lbm3d::k_step_x2<1, true><<<g2, blk, 0, st>>>(p);
            else    lbm3d::k_step_x2<1, false><<<g2, blk, 0, st>>>(p);
        } else {
            if (g1) lbm3d::k_step_x2<0, true><<<g2, blk, 0, st>>>(p);
            else    lbm3d::k_step_x2<0, false><<<g2, blk, 0, st>>>(p);
        }
        c->launches++;
        return;
    }
    switch (mode) {
        case lbm3d::MODE_STEP: launch_step_mode<lbm3d::MODE_STEP>(c, p, grid, blk, st); break;
        case lbm3d::MODE_STEP_STOREF: launch_step_mode<lbm3d::MODE_STEP_STOREF>(c, p, grid, blk, st); break;
        case lbm3d::MODE_STEP_DIFF_STOREF: launch_step_mode<lbm3d::MODE_STEP_DIFF_STOREF>(c, p, grid, blk, st); break;
        default: launch_step_mode<lbm3d::MODE_MATERIALISE>(c, p, grid, blk, st); break;
    }
}

static long long blocks_per_plane(const lbm_ctx *c) {
    dim3 blk, g;
    choose_block(c, &blk, &g);
    return (long long)g.x * g.y;
}

static int ensure_Fbuf(lbm_ctx *c) {
    if (c->Fbuf) return LBM_OK;
    int rc = dev_alloc(c, &c->Fbuf, (long long)lbm3d::Q * c->FS);
    if (rc) return rc;
    c->f_valid = false;
    return LBM_OK;
}

static int ensure_partials(lbm_ctx *c) {
    const long long need = blocks_per_plane(c) * c->nzl;
    if (c->partials && c->npartials >= need) return LBM_OK;
    if (c->partials) { cudaFree(c->partials); c->bytes -= c->npartials * 8; c->partials = nullptr; }
    int rc = dev_alloc(c, &c->partials, need);
    if (rc) return rc;
    c->npartials = need;
    return LBM_OK;
}

// Exchange the face-crossing populations of lattice `L` (just written) with the z-neighbours.
// Up-going populations (cz=+1) of my top plane -> bottom ghost of rank+1; down-going (cz=-1) of my
// bottom plane -> top ghost of rank-1.  Each is one contiguous N^2 block (z is the slowest axis, 3D:48).
static int halo_exchange(lbm_ctx *c, double *L, cudaStream_t st) {
    const int r = c->prm.rank, P = c->prm.nranks;
    const size_t n = (size_t)c->N2;
    NK(g_nccl.GroupStart());
    for (int a = 0; a < 5; a++) {
        if (r + 1 < P) {
            NK(g_nccl.Send(L + (long long)Q_UP[a] * c->PS + (long long)c->nzl * c->N2, n, ncclDouble, r + 1, c->comm, st));
            NK(g_nccl.Recv(L + (long long)Q_DN[a] * c->PS + (long long)(c->nzl + 1) * c->N2, n, ncclDouble, r + 1, c->comm, st));
        }
        if (r > 0) {
            NK(g_nccl.Send(L + (long long)Q_DN[a] * c->PS + (long long)1 * c->N2, n, ncclDouble, r - 1, c->comm, st));
            NK(g_nccl.Recv(L + (long long)Q_UP[a] * c->PS + (long long)0 * c->N2, n, ncclDouble, r - 1, c->comm, st));
        }
    }
    NK(g_nccl.GroupEnd());
    return LBM_OK;
}

static const int GRAPH_LEN = 60;

static void drop_graphs(lbm_ctx *c) {
    for (int k = 0; k < 6; k++)
        if (c->graph_exec[k]) { cudaGraphExecDestroy(c->graph_exec[k]); c->graph_exec[k] = nullptr; }
}

static bool graphs_enabled(const lbm_ctx *c) {
    if (c->prm.dim != 3 || c->prm.nranks != 1) return false;
    if (c->graphs >= 0) return c->graphs != 0;
    return c->FS <= (1LL << 21);   // auto: up to 128^3 cells the step is launch-latency-bound
}

static int run_iteration(lbm_ctx *c, int mode);

// GRAPH_LEN plain iterations as one graph launch (captured on first use of each phase).
static int run_graph_block(lbm_ctx *c) {
    const int phase = (int)(c->steps % 3) * 2 + c->cur;
    if (!c->graph_exec[phase]) {
        const long long steps0 = c->steps, launches0 = c->launches;
        const int cur0 = c->cur;
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(c->s_main, cudaStreamCaptureModeThreadLocal));
        int rc = LBM_OK;
        for (int k = 0; k < GRAPH_LEN && !rc; k++) rc = run_iteration(c, lbm3d::MODE_STEP);
        cudaError_t e = cudaStreamEndCapture(c->s_main, &g);
        c->steps = steps0; c->launches = launches0; c->cur = cur0;   // capturing executed nothing
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&c->graph_exec[phase], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    CK(cudaGraphLaunch(c->graph_exec[phase], c->s_main));
    c->steps += GRAPH_LEN;        // GRAPH_LEN is even and a multiple of 3: cur and the side slot phase are unchanged
    c->launches += GRAPH_LEN;
    c->f_valid = false;
    c->residual_ready = false;
    return LBM_OK;
}

// One iteration in the given mode.  Reads lat[cur], writes lat[cur^1] (unless MATERIALISE).
static int run_iteration(lbm_ctx *c, int mode) {
    dim3 blk, gxy;
    choose_block(c, &blk, &gxy);
    lbm3d::KP p = make_kp(c);
    const bool real = (mode != lbm3d::MODE_MATERIALISE);
    // iteration index t = c->steps (real step) or c->steps - 1 (re-materialising the last one)
    const long long t = real ? c->steps : c->steps - 1;
    const int src_i = real ? c->cur : (c->cur ^ 1);
    kp_set_lattices(c, &p, c->lat[src_i], c->lat[src_i ^ 1]);
    const long long side_n = (long long)c->nzl * c->N;
    p.side_rd = c->side + ((t + 1) % 3) * side_n;  // written by iteration t-2
    p.side_wr = c->side + (t % 3) * side_n;
    const long long bpp = (long long)gxy.x * gxy.y;

    const bool multi = c->prm.nranks > 1;
    if (!multi || !real) {
        if (multi && c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
        p.kl0 = 0; p.kstride = 1; p.partial_base = 0;
        launch_step(c, mode, p, dim3(gxy.x, gxy.y, c->nzl), blk, c->s_main);
    } else {
        // Two streams.  s_halo (high priority): the two slab-boundary planes, then the NCCL exchange
        // of what they produced.  s_main: the interior planes, concurrently (the boundary launch is
        // too small to fill the GPU on its own, and the exchange hides behind the interior).
        //   interior(t) needs boundary(t-1): it reads planes 0 / nzl-1 of the source lattice and
        //                overwrites planes 1 / nzl-2 of the lattice boundary(t-1) was reading;
        //   boundary(t) needs interior(t-1) (same two reasons, mirrored) and exchange(t-1) (ghosts;
        //                stream order on s_halo).
        CK(cudaStreamWaitEvent(c->s_main, c->ev_bnd, 0));   // the record of iteration t-1 (no-op the first time)
        CK(cudaEventRecord(c->ev_main, c->s_main));         // everything queued on s_main so far, incl. interior(t-1)
        CK(cudaStreamWaitEvent(c->s_halo, c->ev_main, 0));
        const int nb = c->nzl >= 2 ? 2 : 1;
        p.kl0 = 0; p.kstride = c->nzl >= 2 ? c->nzl - 1 : 1; p.partial_base = 0;
        if (!c->p2p) {
            launch_step(c, mode, p, dim3(gxy.x, gxy.y, nb), blk, c->s_halo);
            CK(cudaEventRecord(c->ev_bnd, c->s_halo));
            int rc = halo_exchange(c, c->lat[src_i ^ 1], c->s_halo);
            if (rc) return rc;
        } else {
            // Direct NVLink halo: wait until both neighbours finished their boundary kernels of the
            // previous iteration (they filled my ghosts and are done reading the ghosts I overwrite
            // now), then ONE kernel computes the boundary planes, stores the face-crossing
            // populations into the neighbours' ghost planes and publishes this iteration's stamp.
            const long long rel = t - c->p2p_base;
            const int r = c->prm.rank, P = c->prm.nranks;
            lbm3d::k_wait_arrivals<<<1, 1, 0, c->s_halo>>>(c->sig, r > 0 ? rel : 0, r + 1 < P ? rel : 0, c->d_flag,
                                                            20000000000LL);
            c->launches++;
            const int dsti = src_i ^ 1;
            for (int a = 0; a < 5; a++) {
                p.peer_dn[a] = r > 0 ? c->peer_lat[0][dsti] + (long long)Q_DN[a] * c->peer_PS[0] + (long long)(c->peer_nzl[0] + 1) * c->N2
                                     : nullptr;
                p.peer_up[a] = r + 1 < P ? c->peer_lat[1][dsti] + (long long)Q_UP[a] * c->peer_PS[1] : nullptr;
            }
            p.done_count = c->done_count;
            p.sig_dn = r > 0 ? c->peer_sig[0] + 1 : nullptr;      // I am the neighbour "above" rank-1
            p.sig_up = r + 1 < P ? c->peer_sig[1] + 0 : nullptr;  // and the neighbour "below" rank+1
            p.sig_value = rel + 1;
            const dim3 grid(gxy.x, gxy.y, nb);
            switch (mode) {
                case lbm3d::MODE_STEP: launch_step_peer<lbm3d::MODE_STEP>(c, p, grid, blk, c->s_halo); break;
                case lbm3d::MODE_STEP_STOREF: launch_step_peer<lbm3d::MODE_STEP_STOREF>(c, p, grid, blk, c->s_halo); break;
                default: launch_step_peer<lbm3d::MODE_STEP_DIFF_STOREF>(c, p, grid, blk, c->s_halo); break;
            }
            CK(cudaEventRecord(c->ev_bnd, c->s_halo));
        }
        CK(cudaEventRecord(c->ev_halo, c->s_halo));
        c->halo_pending = true;
        if (c->nzl > 2) {
            p.kl0 = 1; p.kstride = 1; p.partial_base = (int)(bpp * nb);
            launch_step(c, mode, p, dim3(gxy.x, gxy.y, c->nzl - 2), blk, c->s_main);
        }
        if (mode == lbm3d::MODE_STEP_DIFF_STOREF) CK(cudaStreamWaitEvent(c->s_main, c->ev_bnd, 0));  // its block partials
    }
    CK(cudaGetLastError());
    if (real) {
        c->cur ^= 1;
        c->steps++;
        c->f_valid = (mode != lbm3d::MODE_STEP);
        c->residual_ready = false;
        if (mode == lbm3d::MODE_STEP_DIFF_STOREF) {
            lbm3d::k_sum_partials<<<1, 1024, 0, c->s_main>>>(c->partials, bpp * c->nzl, c->d_red);
            c->launches++;
            CK(cudaGetLastError());
            c->residual_ready = true;
        }
    } else {
        c->f_valid = true;
    }
    return LBM_OK;
}

// Make Fbuf hold the reference's f2 of the latest iteration.
static int materialise_F(lbm_ctx *c) {
    int rc = ensure_Fbuf(c);
    if (rc) return rc;
    if (c->f_valid) return LBM_OK;
    if (c->steps == 0) {
        const long long n = c->FS;
        lbm3d::k_fill_w<<<(unsigned)((n + 255) / 256), 256, 0, c->s_main>>>(c->Fbuf, c->FS);
        c->launches++;
        CK(cudaGetLastError());
        c->f_valid = true;
        return LBM_OK;
    }
    rc = run_iteration(c, lbm3d::MODE_MATERIALISE);
    if (rc) return rc;
    if (c->p2p) {
        // With the direct halo a neighbour may already be one iteration ahead and would overwrite the
        // ghost planes this kernel reads; a collective point after it keeps everybody behind it
        // (so lbm_get_f / lbm_macrovars on a stale F buffer are collective calls in this mode).
        NK(g_nccl.AllReduce(c->d_red + 1, c->d_red + 1, 1, ncclDouble, ncclSum, c->comm, c->s_main));
    }
    return LBM_OK;
}

static int poll_flag(lbm_ctx *c) {
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    if (c->s_halo) CK(cudaStreamSynchronize(c->s_halo));
    if (flag & 4) return fail(LBM_ERR_CUDA, "direct halo: timed out waiting for a neighbour GPU's arrival flag");
    if (flag & 1) c->unstable = true;
    if (c->unstable) return fail(LBM_ERR_UNSTABLE, "Error: negative feq.  Therefore, unstable.");
    return LBM_OK;
}


// =====================================================================================================
// dim = 2: D2Q9 cavity (2D main.cpp).  Row slabs along y, ghost row below/above, same stream/event
// scheme as the 3D path; exchanged populations are the reference's own (2D:208-270): {2,5,6} upward,
// {4,7,8} downward, N doubles each.
// =====================================================================================================
static const int CX9[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};   // implied by the pulls 2D:292-348
static const int CY9[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const int Q2_UP[3] = {2, 5, 6};
static const int Q2_DN[3] = {4, 7, 8};

static lbm2d::KP2 make_kp2(const lbm_ctx *c, const double *src, double *dst) {
    lbm2d::KP2 p;
    memset(&p, 0, sizeof(p));
    for (int k = 0; k < 9; k++) {
        const long long shift = (long long)CX9[k] + (long long)CY9[k] * c->N;
        p.srcq[k] = src ? src + (long long)k * c->PS - shift : nullptr;
        p.own[k] = src ? src + (long long)k * c->PS : nullptr;
        p.dstq[k] = dst ? dst + (long long)k * c->PS : nullptr;
    }
    p.u = c->u2; p.v = c->v2; p.rho = c->rho2;
    p.partials = c->partials;
    p.N = c->N; p.M = c->M2; p.nyl = c->nyl; p.j0 = c->j0;
    p.jl0 = 0; p.jstride = 1; p.partial_base = 0;
    p.omega = c->OMEGA; p.homega = 0.5 * c->OMEGA; p.homegam = 0.5 * c->OMEGAm;  // 2D:365-366
    p.lid6 = c->lid6;
    return p;
}

static int block2d(const lbm_ctx *c) {
    if (c->block_x > 0) return c->block_x;
    int bx = ((c->N + 31) / 32) * 32;
    if (bx > 256) bx = 256;  // 256 measured best at 8192^2 (profiles/r01_sweep2d.log)
    return bx;
}

static int halo_exchange2d(lbm_ctx *c, double *L, cudaStream_t st) {
    const int r = c->prm.rank, P = c->prm.nranks;
    const size_t n = (size_t)c->N;
    NK(g_nccl.GroupStart());
    for (int a = 0; a < 3; a++) {
        if (r + 1 < P) {
            NK(g_nccl.Send(L + (long long)Q2_UP[a] * c->PS + (long long)c->nyl * c->N, n, ncclDouble, r + 1, c->comm, st));
            NK(g_nccl.Recv(L + (long long)Q2_DN[a] * c->PS + (long long)(c->nyl + 1) * c->N, n, ncclDouble, r + 1, c->comm, st));
        }
        if (r > 0) {
            NK(g_nccl.Send(L + (long long)Q2_DN[a] * c->PS + (long long)c->N, n, ncclDouble, r - 1, c->comm, st));
            NK(g_nccl.Recv(L + (long long)Q2_UP[a] * c->PS, n, ncclDouble, r - 1, c->comm, st));
        }
    }
    NK(g_nccl.GroupEnd());
    return LBM_OK;
}

static int create2d(lbm_ctx *c) {
    const lbm_params &p = c->prm;
    c->N = p.N;
    c->M2 = p.M;
    c->j0 = (int)((long long)p.rank * p.M / p.nranks);
    c->j1 = (int)((long long)(p.rank + 1) * p.M / p.nranks);
    c->nyl = c->j1 - c->j0;
    c->k0 = c->j0; c->k1 = c->j1;
    if (c->nyl < 1) return fail(LBM_ERR_INVALID, "rank %d of %d owns no row of M=%d", p.rank, p.nranks, p.M);
    c->N2 = p.N;                                 // "plane" = one row
    c->PS = (long long)(c->nyl + 2) * c->N;
    c->FS = (long long)c->nyl * c->N;
    if (c->PS >= (1LL << 31)) return fail(LBM_ERR_INVALID, "more than 2^31 cells per GPU");
    // 2D:22-26, evaluated in the same order
    c->NU = p.UMAX * p.N / p.Re;
    c->TAU = c->NU * 3.0 + 0.5;
    c->OMEGA = 1.0 / c->TAU;
    c->OMEGAm = 1.0 / (0.5 + (p.MAGIC / (c->TAU - 0.5)));
    c->OmOMEGA = 1.0 - c->OMEGA;
    c->gammainv = 1.0;
    c->ulid = p.UMAX; c->vlid = 0.0; c->wlid = 0.0;
    c->lid6 = 6.0 * LBM2_WD * p.UMAX;            // 2D:354

    int rc;
    c->pad = ((c->N + 1 + 15) / 16) * 16;
    const long long tot = 9LL * c->PS + 2 * c->pad;
    for (int l = 0; l < 2; l++) {
        if ((rc = dev_alloc(c, &c->lat_alloc[l], tot))) return rc;
        CK(cudaMemsetAsync(c->lat_alloc[l], 0, (size_t)tot * sizeof(double), c->s_main));  // dist = 0, 2D:127,138
        c->lat[l] = c->lat_alloc[l] + c->pad;
    }
    if ((rc = dev_alloc(c, &c->u2, c->FS))) return rc;
    if ((rc = dev_alloc(c, &c->v2, c->FS))) return rc;
    if ((rc = dev_alloc(c, &c->rho2, c->FS))) return rc;
    if ((rc = dev_alloc(c, &c->d_red, 2))) return rc;
    if ((rc = dev_alloc(c, &c->d_flag, 1))) return rc;
    CK(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->s_main));
    CK(cudaMemsetAsync(c->d_red, 0, 2 * sizeof(double), c->s_main));
    CK(cudaMemsetAsync(c->u2, 0, (size_t)c->FS * sizeof(double), c->s_main));   // u = v = 0, rho = 1, 2D:128-133
    CK(cudaMemsetAsync(c->v2, 0, (size_t)c->FS * sizeof(double), c->s_main));
    lbm3d::k_fill<<<(unsigned)((c->FS + 255) / 256), 256, 0, c->s_main>>>(c->rho2, c->FS, 1.0);
    c->launches++;
    CK(cudaMallocHost((void **)&c->h_pin, 4 * sizeof(double)));
    const int bx = block2d(c);
    c->npartials = (long long)((c->N + bx - 1) / bx) * c->nyl;
    if ((rc = dev_alloc(c, &c->partials, c->npartials))) return rc;
    c->cur = 0;
    c->steps = 0;
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

static int step2d(lbm_ctx *c, long long n) {
    const int bx = block2d(c);
    const unsigned gx = (unsigned)((c->N + bx - 1) / bx);
    const bool multi = c->prm.nranks > 1;
    for (long long t = 0; t < n; t++) {
        lbm2d::KP2 p = make_kp2(c, c->lat[c->cur], c->lat[c->cur ^ 1]);
        if (multi && c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
        if (!multi) {
            lbm2d::k2_step<<<dim3(gx, c->nyl), bx, 0, c->s_main>>>(p);
            c->launches++;
        } else {
            const int nb = c->nyl >= 2 ? 2 : 1;
            p.jl0 = 0; p.jstride = c->nyl >= 2 ? c->nyl - 1 : 1;
            lbm2d::k2_step<<<dim3(gx, nb), bx, 0, c->s_main>>>(p);
            c->launches++;
            CK(cudaEventRecord(c->ev_bnd, c->s_main));
            CK(cudaStreamWaitEvent(c->s_halo, c->ev_bnd, 0));
            int rc = halo_exchange2d(c, c->lat[c->cur ^ 1], c->s_halo);
            if (rc) return rc;
            CK(cudaEventRecord(c->ev_halo, c->s_halo));
            c->halo_pending = true;
            if (c->nyl > 2) {
                p.jl0 = 1; p.jstride = 1;
                lbm2d::k2_step<<<dim3(gx, c->nyl - 2), bx, 0, c->s_main>>>(p);
                c->launches++;
            }
        }
        c->cur ^= 1;
        c->steps++;
    }
    CK(cudaGetLastError());
    return LBM_OK;
}

// macroVars + convergence (2D:419-478): change of (u, v, rho) since the previous call
static int residual2d(lbm_ctx *c, double *out) {
    const int bx = block2d(c);
    const unsigned gx = (unsigned)((c->N + bx - 1) / bx);
    if ((long long)gx * c->nyl > c->npartials) {
        cudaFree(c->partials); c->bytes -= c->npartials * 8; c->partials = nullptr;
        c->npartials = (long long)gx * c->nyl;
        int rc = dev_alloc(c, &c->partials, c->npartials);
        if (rc) return rc;
    }
    lbm2d::KP2 p = make_kp2(c, c->lat[c->cur], nullptr);
    lbm2d::k2_macro<true><<<dim3(gx, c->nyl), bx, 0, c->s_main>>>(p);
    lbm3d::k_sum_partials<<<1, 1024, 0, c->s_main>>>(c->partials, (long long)gx * c->nyl, c->d_red);
    c->launches += 2;
    CK(cudaGetLastError());
    const double *src = c->d_red;
    if (c->prm.nranks > 1) {
        NK(g_nccl.AllReduce(c->d_red, c->d_red + 1, 1, ncclDouble, ncclSum, c->comm, c->s_main));  // 2D:457
        src = c->d_red + 1;
    }
    CK(cudaMemcpyAsync(c->h_pin, src, sizeof(double), cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    *out = sqrt(c->h_pin[0]);  // 2D:460
    return LBM_OK;
}

// dist in the reference's layout LU3(i,j,k) = 9N*j + N*k + i (2D:276-278)
static int copy_dist2d(lbm_ctx *c, double *host, bool to_host) {
    const long long hj0 = c->host_local ? 0 : c->j0;
    for (int k = 0; k < 9; k++) {
        double *h = host + hj0 * 9 * c->N + (long long)k * c->N;
        double *d = c->lat[c->cur] + (long long)k * c->PS + c->N;
        if (to_host)
            CK(cudaMemcpy2DAsync(h, (size_t)9 * c->N * 8, d, (size_t)c->N * 8, (size_t)c->N * 8, (size_t)c->nyl,
                                 cudaMemcpyDeviceToHost, c->s_main));
        else
            CK(cudaMemcpy2DAsync(d, (size_t)c->N * 8, h, (size_t)9 * c->N * 8, (size_t)c->N * 8, (size_t)c->nyl,
                                 cudaMemcpyHostToDevice, c->s_main));
    }
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

static int macrovars2d(lbm_ctx *c, double *u, double *v, double *rho, double *vmag) {
    int rc;
    for (int a = 0; a < 3; a++)
        if (!c->out2[a]) { if ((rc = dev_alloc(c, &c->out2[a], c->FS))) return rc; }
    const int bx = block2d(c);
    const unsigned gx = (unsigned)((c->N + bx - 1) / bx);
    lbm2d::KP2 p = make_kp2(c, c->lat[c->cur], nullptr);
    p.u = c->out2[0]; p.v = c->out2[1]; p.rho = c->out2[2];
    lbm2d::k2_macro<false><<<dim3(gx, c->nyl), bx, 0, c->s_main>>>(p);
    c->launches++;
    CK(cudaGetLastError());
    const long long off = c->host_local ? 0 : (long long)c->j0 * c->N;
    double *host[3] = {u, v, rho};
    for (int a = 0; a < 3; a++)
        if (host[a])
            CK(cudaMemcpyAsync(host[a] + off, c->out2[a], (size_t)c->FS * 8, cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    if (vmag && u && v)
        for (long long n = 0; n < c->FS; n++) vmag[off + n] = sqrt(u[off + n] * u[off + n] + v[off + n] * v[off + n]);
    return LBM_OK;
}

static cudaEvent_t span_event(lbm_ctx *c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Fold finished spans into plain_ms / plain_iters (wait == true: synchronise on the open ones too).
static void resolve_spans(lbm_ctx *c, bool wait) {
    size_t keep = 0;
    for (size_t k = 0; k < c->spans.size(); k++) {
        lbm_ctx::Span &sp = c->spans[k];
        cudaError_t e = wait ? cudaEventSynchronize(sp.b) : cudaEventQuery(sp.b);
        if (e == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) { c->plain_ms += ms; c->plain_iters += sp.iters; }
            c->ev_pool.push_back(sp.a);
            c->ev_pool.push_back(sp.b);
        } else {
            c->spans[keep++] = sp;
        }
    }
    c->spans.resize(keep);
    cudaGetLastError();  // cudaErrorNotReady from the queries is not an error
}

// ---- C ABI ---------------------------------------------------------------------------------------
extern "C" {

const char *lbm_last_error(void) { return g_err; }

const char *lbm_version(void) { return "lbm_b200 0.1 (sm_100a, fp64, -fmad=false, D3Q19 pull/two-lattice)"; }

void lbm_default_params(lbm_params *p) {
    memset(p, 0, sizeof(*p));
    p->dim = 3; p->N = 50; p->M = 0; p->Re = 100; p->UMAX = 0.1;   // 3D:7-9
    p->METHOD = 1; p->GAMMA = 1.0; p->MAGIC = 0.25; p->DIAG = 0;   // 3D:12-14, 22
    p->rank = 0; p->nranks = 1; p->device = -1; p->track_residual = 1; p->stream = nullptr;
}

int lbm_nccl_unique_id(unsigned char out[LBM_NCCL_ID_BYTES]) {
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == LBM_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(out, &id, LBM_NCCL_ID_BYTES);
    return LBM_OK;
}

int lbm_destroy(lbm_ctx *c) {
    if (!c) return LBM_OK;
    cudaSetDevice(c->dev);
    if (c->p2p && c->s_halo && c->steps > c->p2p_base) {
        // the neighbours' last boundary kernels store into my ghost planes: let them land first
        const long long rel = c->steps - c->p2p_base;
        lbm3d::k_wait_arrivals<<<1, 1, 0, c->s_halo>>>(c->sig, c->prm.rank > 0 ? rel : 0,
                                                        c->prm.rank + 1 < c->prm.nranks ? rel : 0, c->d_flag, 4000000000LL);
    }
    if (c->s_main) cudaStreamSynchronize(c->s_main);
    if (c->s_halo) cudaStreamSynchronize(c->s_halo);
    for (int side = 0; side < 2; side++)
        for (int a = 0; a < 3; a++)
            if (c->peer_ipc[side][a]) cudaIpcCloseMemHandle(c->peer_ipc[side][a]);
    cudaFree(c->sig); cudaFree(c->done_count);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->lat_alloc[0]); cudaFree(c->lat_alloc[1]); cudaFree(c->Fbuf); cudaFree(c->side);
    cudaFree(c->partials); cudaFree(c->d_red); cudaFree(c->d_flag);
    for (int a = 0; a < 5; a++) cudaFree(c->d_out[a]);
    cudaFree(c->u2); cudaFree(c->v2); cudaFree(c->rho2);
    for (int a = 0; a < 3; a++) cudaFree(c->out2[a]);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    drop_graphs(c);
    for (auto &sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->s_halo) cudaStreamDestroy(c->s_halo);
    if (c->own_main && c->s_main) cudaStreamDestroy(c->s_main);
    delete c;
    return LBM_OK;
}

static int create3d(lbm_ctx *c) {
    const lbm_params &p = c->prm;
    c->N = p.N;
    c->N2 = (long long)p.N * p.N;
    c->k0 = (int)((long long)p.rank * p.N / p.nranks);
    c->k1 = (int)((long long)(p.rank + 1) * p.N / p.nranks);
    c->nzl = c->k1 - c->k0;
    if (c->nzl < 1) return fail(LBM_ERR_INVALID, "rank %d of %d owns no plane of N=%d", p.rank, p.nranks, p.N);
    c->PS = (long long)(c->nzl + 2) * c->N2;
    c->FS = (long long)c->nzl * c->N2;
    if (c->PS >= (1LL << 31)) return fail(LBM_ERR_INVALID, "more than 2^31 cells per GPU");

    // derived constants in the reference's macro order (3D:19-23, 30, 40) so the doubles are
    // bit-identical to gcc's constant folding
    const int Re = (int)p.Re;
    c->NU = (p.UMAX) * (double)(p.N + 1.0) / (double)(Re);
    c->TAU = (c->NU) * (3.0 / (double)p.GAMMA) + 0.5;
    c->OMEGA = 1.0 / (double)(c->TAU);
    c->OMEGAm = 1.0 / (0.5 + (p.MAGIC / ((1.0 / c->OMEGA) - 0.5)));
    c->OmOMEGA = 1.0 - c->OMEGA;
    c->gammainv = 1.0 / p.GAMMA;
    if (p.DIAG == 1) {  // 3D:183-191
        c->ulid = (double)p.UMAX / sqrt(2.0);
        c->wlid = c->ulid;
    } else {
        c->ulid = p.UMAX;
        c->wlid = 0.0;
    }
    c->vlid = 0.0;

    int rc;
    c->pad = ((c->N2 + c->N + 1 + 15) / 16) * 16;  // shifted pulls of box-boundary cells stay inside the allocation
    for (int l = 0; l < 2; l++) {
        if ((rc = dev_alloc(c, &c->lat_alloc[l], (long long)lbm3d::Q * c->PS + 2 * c->pad))) return rc;
        CK(cudaMemsetAsync(c->lat_alloc[l], 0, (size_t)c->pad * sizeof(double), c->s_main));
        CK(cudaMemsetAsync(c->lat_alloc[l] + c->pad + (long long)lbm3d::Q * c->PS, 0, (size_t)c->pad * sizeof(double), c->s_main));
        c->lat[l] = c->lat_alloc[l] + c->pad;
    }
    const long long side_n = (long long)c->nzl * c->N;
    if ((rc = dev_alloc(c, &c->side, 3 * side_n))) return rc;
    if ((rc = dev_alloc(c, &c->d_red, 2))) return rc;
    if ((rc = dev_alloc(c, &c->d_flag, 1))) return rc;
    CK(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->s_main));
    CK(cudaMemsetAsync(c->d_red, 0, 2 * sizeof(double), c->s_main));
    CK(cudaMallocHost((void **)&c->h_pin, 4 * sizeof(double)));

    // stale-f14 history starts at the initial value w[14] = 1/36 (3D:192-205)
    lbm3d::k_fill<<<(unsigned)((3 * side_n + 255) / 256), 256, 0, c->s_main>>>(c->side, 3 * side_n, LBM_W2);
    c->launches++;

    // S(-1) = collide(f = w) in both lattices incl. ghost planes (colliding the initial equilibrium
    // is NOT a no-op in fp64: sum of w[q] != 1)
    dim3 blk, gxy;
    choose_block(c, &blk, &gxy);
    for (int l = 0; l < 2; l++) {
        lbm3d::KP kp = make_kp(c);
        kp_set_lattices(c, &kp, nullptr, c->lat[l]);
        kp.Fbuf = nullptr;
        dim3 grid(gxy.x, gxy.y, c->nzl + 2);
        if (p.METHOD == 1) lbm3d::k_collide_cells<1><<<grid, blk, 0, c->s_main>>>(kp, 1);
        else               lbm3d::k_collide_cells<0><<<grid, blk, 0, c->s_main>>>(kp, 1);
        c->launches++;
    }
    CK(cudaGetLastError());
    c->cur = 0;
    c->steps = 0;
    if (p.track_residual) {
        if ((rc = materialise_F(c))) return rc;  // Fbuf = w: the reference's f1 before iteration 0 (3D:96)
        if ((rc = ensure_partials(c))) return rc;
    }
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

int lbm_create(const lbm_params *p, lbm_ctx **out) {
    if (!p || !out) return fail(LBM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (p->dim != 3 && p->dim != 2) return fail(LBM_ERR_INVALID, "dim=%d (3 = D3Q19 cube, 2 = D2Q9 N x M)", p->dim);
    if (p->N < 3) return fail(LBM_ERR_INVALID, "N=%d: need at least 3 grid points per side", p->N);
    if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) return fail(LBM_ERR_INVALID, "bad rank/nranks");
    if (p->dim == 3) {
        if (p->METHOD != 0 && p->METHOD != 1) return fail(LBM_ERR_INVALID, "METHOD=%d (0 = BGK, 1 = TRT)", p->METHOD);
        if (!(p->Re > 0) || p->Re != floor(p->Re)) return fail(LBM_ERR_INVALID, "Re must be a positive integer (3D:7)");
        if (!(p->GAMMA > 0)) return fail(LBM_ERR_INVALID, "GAMMA must be positive");
        if (p->nranks > p->N) return fail(LBM_ERR_INVALID, "more ranks than z planes");
    } else {
        if (p->M < 3) return fail(LBM_ERR_INVALID, "M=%d: need at least 3 rows", p->M);
        if (p->METHOD != 1) return fail(LBM_ERR_INVALID, "dim=2 is TRT only (2D:367-413): METHOD must be 1");
        if (!(p->Re > 0)) return fail(LBM_ERR_INVALID, "Re must be positive");
        if (p->nranks > p->M) return fail(LBM_ERR_INVALID, "more ranks than rows");
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(LBM_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    lbm_ctx *c = new lbm_ctx();
    c->prm = *p;
    if (p->device >= 0) {
        if (p->device >= ndev) { delete c; return fail(LBM_ERR_INVALID, "device %d of %d", p->device, ndev); }
        c->dev = p->device;
    } else {
        cudaGetDevice(&c->dev);
    }
    int rc = LBM_OK;
#define CKD(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            rc = fail(LBM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                              \
            lbm_destroy(c);                                                                                \
            return rc;                                                                                     \
        }                                                                                                  \
    } while (0)
    CKD(cudaSetDevice(c->dev));
    if (p->stream) {
        c->s_main = (cudaStream_t)p->stream;
    } else {
        CKD(cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking));
        c->own_main = true;
    }
    CKD(cudaEventCreate(&c->ev_t0));
    CKD(cudaEventCreate(&c->ev_t1));
    if (p->nranks > 1) {
        int lo, hi;
        CKD(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CKD(cudaStreamCreateWithPriority(&c->s_halo, cudaStreamNonBlocking, hi));
        CKD(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
        if ((rc = nccl_load())) { lbm_destroy(c); return rc; }
        ncclUniqueId id;
        memcpy(&id, p->nccl_id, LBM_NCCL_ID_BYTES);
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, p->nranks, id, p->rank);
        if (r != ncclSuccess) {
            rc = fail(LBM_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
            c->comm = nullptr;
            lbm_destroy(c);
            return rc;
        }
    }
#undef CKD
    rc = (p->dim == 3) ? create3d(c) : create2d(c);
    if (rc) { lbm_destroy(c); return rc; }
    *out = c;
    return LBM_OK;
}


int lbm_peer_export(lbm_ctx *c, unsigned char blob[LBM_PEER_BLOB_BYTES]) {
    if (!c || !blob) return fail(LBM_ERR_INVALID, "null argument");
    if (c->prm.dim != 3 || c->prm.nranks < 2) return fail(LBM_ERR_INVALID, "the direct halo needs dim=3 and nranks > 1");
    CK(cudaSetDevice(c->dev));
    int rc;
    if (!c->sig) {
        if ((rc = dev_alloc(c, &c->sig, 2))) return rc;
        if ((rc = dev_alloc(c, &c->done_count, 1))) return rc;
        CK(cudaMemset(c->sig, 0, 2 * sizeof(long long)));
        CK(cudaMemset(c->done_count, 0, sizeof(unsigned int)));
    }
    PeerBlob b;
    memset(&b, 0, sizeof b);
    b.magic = 0x4c424d50; b.pid = (int)getpid(); b.dev = c->dev; b.nzl = c->nzl; b.PS = c->PS; b.pad = c->pad;
    CK(cudaIpcGetMemHandle(&b.lat[0], c->lat_alloc[0]));
    CK(cudaIpcGetMemHandle(&b.lat[1], c->lat_alloc[1]));
    CK(cudaIpcGetMemHandle(&b.sig, c->sig));
    b.raw_lat[0] = c->lat_alloc[0]; b.raw_lat[1] = c->lat_alloc[1]; b.raw_sig = c->sig;
    memset(blob, 0, LBM_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof b);
    return LBM_OK;
}

static int peer_open(lbm_ctx *c, int side, const unsigned char *blob) {
    PeerBlob b;
    memcpy(&b, blob, sizeof b);
    if (b.magic != 0x4c424d50) return fail(LBM_ERR_INVALID, "bad peer blob");
    void *lat0 = nullptr, *lat1 = nullptr, *sg = nullptr;
    if (b.pid == (int)getpid()) {  // same process (one thread per GPU): plain peer access
        if (b.dev != c->dev) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->dev, b.dev));
            if (!can) return fail(LBM_ERR_CUDA, "device %d cannot access device %d", c->dev, b.dev);
            cudaError_t e = cudaDeviceEnablePeerAccess(b.dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(LBM_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        lat0 = b.raw_lat[0]; lat1 = b.raw_lat[1]; sg = b.raw_sig;
    } else {
        CK(cudaIpcOpenMemHandle(&lat0, b.lat[0], cudaIpcMemLazyEnablePeerAccess));
        CK(cudaIpcOpenMemHandle(&lat1, b.lat[1], cudaIpcMemLazyEnablePeerAccess));
        CK(cudaIpcOpenMemHandle(&sg, b.sig, cudaIpcMemLazyEnablePeerAccess));
        c->peer_ipc[side][0] = lat0; c->peer_ipc[side][1] = lat1; c->peer_ipc[side][2] = sg;
    }
    c->peer_lat[side][0] = (double *)lat0 + b.pad;
    c->peer_lat[side][1] = (double *)lat1 + b.pad;
    c->peer_sig[side] = (long long *)sg;
    c->peer_nzl[side] = b.nzl;
    c->peer_PS[side] = b.PS;
    return LBM_OK;
}

int lbm_peer_connect(lbm_ctx *c, const unsigned char *below_blob, const unsigned char *above_blob) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (c->prm.dim != 3 || c->prm.nranks < 2) return fail(LBM_ERR_INVALID, "the direct halo needs dim=3 and nranks > 1");
    if (!c->sig) return fail(LBM_ERR_INVALID, "call lbm_peer_export first");
    const int r = c->prm.rank, P = c->prm.nranks;
    if ((r > 0 && !below_blob) || (r + 1 < P && !above_blob)) return fail(LBM_ERR_INVALID, "missing neighbour blob");
    CK(cudaSetDevice(c->dev));
    int rc = poll_flag(c);  // drains both streams
    if (rc) return rc;
    if (r > 0 && (rc = peer_open(c, 0, below_blob))) return rc;
    if (r + 1 < P && (rc = peer_open(c, 1, above_blob))) return rc;
    CK(cudaMemset(c->sig, 0, 2 * sizeof(long long)));
    c->p2p_base = c->steps;
    c->p2p = true;
    return LBM_OK;
}

int lbm_step(lbm_ctx *c, long long n) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (n < 0) return fail(LBM_ERR_INVALID, "negative step count");
    if (c->unstable) return fail(LBM_ERR_UNSTABLE, "Error: negative feq.  Therefore, unstable.");
    CK(cudaSetDevice(c->dev));
    if (c->prm.dim == 2) return step2d(c, n);
    const bool track = c->prm.track_residual != 0;
    int rc;
    if (track && n > 0) {
        if ((rc = ensure_Fbuf(c))) return rc;
        if ((rc = ensure_partials(c))) return rc;
        if (n == 1 && !c->f_valid) { if ((rc = materialise_F(c))) return rc; }
    }
    const long long n_plain = track ? n - 2 : n;  // the leading iterations run the plain step kernel
    lbm_ctx::Span sp{nullptr, nullptr, 0};
    if (n_plain > 0) {
        if (c->spans.size() > 256) resolve_spans(c, false);
        sp.a = span_event(c);
        sp.b = span_event(c);
        sp.iters = n_plain;
        CK(cudaEventRecord(sp.a, c->s_main));
    }
    const bool use_graphs = graphs_enabled(c);
    for (long long t = 0; t < n; t++) {
        int mode = lbm3d::MODE_STEP;
        if (track) {
            if (t == n - 1) mode = lbm3d::MODE_STEP_DIFF_STOREF;     // F(t) vs F(t-1): the reference's diff(f2, f1)
            else if (t == n - 2) mode = lbm3d::MODE_STEP_STOREF;     // leaves F(t-1) for the last iteration
        }
        if (use_graphs && t + GRAPH_LEN <= n_plain) {
            if ((rc = run_graph_block(c))) return rc;
            t += GRAPH_LEN - 1;
        } else if ((rc = run_iteration(c, mode))) {
            return rc;
        }
        if (t == n_plain - 1) {
            CK(cudaEventRecord(sp.b, c->s_main));
            c->spans.push_back(sp);
        }
    }
    return LBM_OK;
}

int lbm_timer_start(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->dev));
    CK(cudaStreamSynchronize(c->s_main));
    if (c->s_halo) CK(cudaStreamSynchronize(c->s_halo));
    CK(cudaEventRecord(c->ev_t0, c->s_main));
    return LBM_OK;
}

int lbm_timer_stop(lbm_ctx *c, float *ms) {
    if (!c || !ms) return fail(LBM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->dev));
    if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
    CK(cudaEventRecord(c->ev_t1, c->s_main));
    CK(cudaEventSynchronize(c->ev_t1));
    CK(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return LBM_OK;
}

int lbm_step_timed(lbm_ctx *c, long long n, float *ms) {
    int rc = lbm_timer_start(c);
    if (rc) return rc;
    if ((rc = lbm_step(c, n))) return rc;
    return lbm_timer_stop(c, ms);
}

int lbm_residual(lbm_ctx *c, double *out) {
    if (!c || !out) return fail(LBM_ERR_INVALID, "null argument");
    if (c->prm.dim == 2) { CK(cudaSetDevice(c->dev)); return residual2d(c, out); }
    if (!c->prm.track_residual) return fail(LBM_ERR_INVALID, "lbm_residual needs track_residual=1");
    if (!c->residual_ready) return fail(LBM_ERR_INVALID, "lbm_residual: no iteration since creation / lbm_set_f");
    CK(cudaSetDevice(c->dev));
    const double *src = c->d_red;
    if (c->prm.nranks > 1) {
        NK(g_nccl.AllReduce(c->d_red, c->d_red + 1, 1, ncclDouble, ncclSum, c->comm, c->s_main));
        src = c->d_red + 1;
    }
    CK(cudaMemcpyAsync(c->h_pin, src, sizeof(double), cudaMemcpyDeviceToHost, c->s_main));
    int rc = poll_flag(c);  // synchronises
    *out = sqrt(c->h_pin[0]);  // 3D:684
    return rc;
}

int lbm_sync(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->dev));
    return poll_flag(c);
}

int lbm_get_f(lbm_ctx *c, double *host_full) {
    if (!c || !host_full) return fail(LBM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->dev));
    if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
    if (c->prm.dim == 2) return copy_dist2d(c, host_full, true);
    int rc = materialise_F(c);
    if (rc) return rc;
    const long long N3 = c->host_local ? c->FS : c->N2 * c->N;
    const long long hk0 = c->host_local ? 0 : c->k0;
    for (int q = 0; q < lbm3d::Q; q++)
        CK(cudaMemcpyAsync(host_full + (long long)q * N3 + hk0 * c->N2, c->Fbuf + (long long)q * c->FS,
                           (size_t)c->FS * sizeof(double), cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

int lbm_set_f(lbm_ctx *c, const double *host_f, const double *host_f_prev) {
    if (!c || !host_f) return fail(LBM_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->dev));
    if (c->prm.dim == 2) {
        if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
        int rc2 = copy_dist2d(c, const_cast<double *>(host_f), false);
        if (rc2) return rc2;
        if (c->prm.nranks > 1) { if ((rc2 = halo_exchange2d(c, c->lat[c->cur], c->s_main))) return rc2; CK(cudaStreamSynchronize(c->s_main)); }
        return LBM_OK;
    }
    int rc = ensure_Fbuf(c);
    if (rc) return rc;
    if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
    const long long N3 = c->host_local ? c->FS : c->N2 * c->N;
    const long long hk0 = c->host_local ? 0 : c->k0;
    for (int q = 0; q < lbm3d::Q; q++)
        CK(cudaMemcpyAsync(c->Fbuf + (long long)q * c->FS, host_f + (long long)q * N3 + hk0 * c->N2,
                           (size_t)c->FS * sizeof(double), cudaMemcpyHostToDevice, c->s_main));
    // stale f14 of the x = N-1 face: the next iteration (index `steps`) reads slot (steps+1)%3.
    // Taken from the reference's other buffer if given, else from the current one.
    {
        const double *srcf = host_f_prev ? host_f_prev : host_f;
        const long long side_n = (long long)c->nzl * c->N;
        double *tmp = (double *)malloc((size_t)side_n * sizeof(double));
        if (!tmp) return fail(LBM_ERR_NOMEM, "malloc");
        for (int kl = 0; kl < c->nzl; kl++)
            for (int j = 0; j < c->N; j++)
                tmp[(long long)kl * c->N + j] = srcf[14LL * N3 + (hk0 + kl) * c->N2 + (long long)j * c->N + (c->N - 1)];
        // slot read by iteration t = steps is (t+1)%3; the one read by t+1 is (t+2)%3 and must be
        // this state's own post-BC f14, which is in host_f.
        cudaError_t e = cudaMemcpyAsync(c->side + ((c->steps + 1) % 3) * side_n, tmp, (size_t)side_n * sizeof(double),
                                        cudaMemcpyHostToDevice, c->s_main);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_main);
        for (int kl = 0; kl < c->nzl; kl++)
            for (int j = 0; j < c->N; j++)
                tmp[(long long)kl * c->N + j] = host_f[14LL * N3 + (hk0 + kl) * c->N2 + (long long)j * c->N + (c->N - 1)];
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(c->side + ((c->steps + 2) % 3) * side_n, tmp, (size_t)side_n * sizeof(double),
                                cudaMemcpyHostToDevice, c->s_main);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_main);
        free(tmp);
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "side upload: %s", cudaGetErrorString(e));
    }
    // S = collide(F) into lat[cur]
    dim3 blk, gxy;
    choose_block(c, &blk, &gxy);
    lbm3d::KP kp = make_kp(c);
    kp_set_lattices(c, &kp, nullptr, c->lat[c->cur]);
    dim3 grid(gxy.x, gxy.y, c->nzl);
    if (c->prm.METHOD == 1) lbm3d::k_collide_cells<1><<<grid, blk, 0, c->s_main>>>(kp, 0);
    else                    lbm3d::k_collide_cells<0><<<grid, blk, 0, c->s_main>>>(kp, 0);
    c->launches++;
    CK(cudaGetLastError());
    if (c->prm.nranks > 1) {
        rc = halo_exchange(c, c->lat[c->cur], c->s_main);
        if (rc) return rc;
    }
    c->f_valid = true;
    c->residual_ready = false;
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

int lbm_macrovars(lbm_ctx *c, double *u1, double *u2, double *u3, double *rho, double *vmag) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->dev));
    if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
    if (c->prm.dim == 2) return macrovars2d(c, u1, u2, rho, vmag);
    int rc = materialise_F(c);
    if (rc) return rc;
    double *host[5] = {u1, u2, u3, rho, vmag};
    for (int a = 0; a < 5; a++)
        if (!c->d_out[a]) { if ((rc = dev_alloc(c, &c->d_out[a], c->FS))) return rc; }
    const long long n = c->FS;
    lbm3d::k_macrovars<<<(unsigned)((n + 255) / 256), 256, 0, c->s_main>>>(c->Fbuf, c->FS, c->d_out[0], c->d_out[1],
                                                                          c->d_out[2], c->d_out[3], c->d_out[4]);
    c->launches++;
    CK(cudaGetLastError());
    for (int a = 0; a < 5; a++)
        if (host[a])
            CK(cudaMemcpyAsync(host[a] + (c->host_local ? 0LL : (long long)c->k0 * c->N2), c->d_out[a], (size_t)n * sizeof(double),
                               cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

int lbm_get_info(lbm_ctx *c, lbm_info *o) {
    if (!c || !o) return fail(LBM_ERR_INVALID, "null argument");
    memset(o, 0, sizeof(*o));
    o->NU = c->NU; o->TAU = c->TAU; o->OMEGA = c->OMEGA; o->OMEGAm = c->OMEGAm; o->OmOMEGA = c->OmOMEGA;
    o->ulid = c->ulid; o->vlid = c->vlid; o->wlid = c->wlid;
    o->Q = c->prm.dim == 2 ? lbm2d::Q : lbm3d::Q; o->k0 = c->k0; o->k1 = c->k1;
    o->cells_local = c->FS; o->cells_global = c->prm.dim == 2 ? (long long)c->N * c->M2 : c->N2 * c->N;
    o->steps_done = c->steps; o->kernel_launches = c->launches; o->device_bytes = c->bytes;
    cudaSetDevice(c->dev);
    resolve_spans(c, true);
    o->plain_step_ms = c->plain_ms; o->plain_step_iterations = c->plain_iters;
    return LBM_OK;
}

int lbm_set_option(lbm_ctx *c, const char *key, long long value) {
    if (!c || !key) return fail(LBM_ERR_INVALID, "null argument");
    drop_graphs(c);  // captured launches bake in the block shape and kernel variant
    if (!strcmp(key, "graphs")) {
        c->graphs = value < 0 ? -1 : (value != 0);
        return LBM_OK;
    }
    if (!strcmp(key, "block_x")) {
        const int lim = c->prm.dim == 2 ? 256 : 128;
        if (value < 0 || value > lim || (value % 32)) return fail(LBM_ERR_INVALID, "block_x must be a multiple of 32 <= %d", lim);
        c->block_x = (int)value;
    } else if (!strcmp(key, "block_y")) {
        if (value < 0 || value > 256) return fail(LBM_ERR_INVALID, "block_y out of range");
        c->block_y = (int)value;
    } else if (!strcmp(key, "vec2")) {
        c->vec2 = value != 0;
        return LBM_OK;
    } else if (!strcmp(key, "host_local_layout")) {
        c->host_local = value != 0;
        return LBM_OK;
    } else if (!strcmp(key, "track_residual")) {
        c->prm.track_residual = value != 0;
        c->residual_ready = false;
        return LBM_OK;
    } else {
        return fail(LBM_ERR_INVALID, "unknown option %s", key);
    }
    if (c->prm.dim == 3 && c->block_x > 0 && c->block_y > 0 && c->block_x * c->block_y > 128)
        return fail(LBM_ERR_INVALID, "block_x*block_y > 128");
    // partial-sum buffer depends on the block shape
    if (c->prm.dim == 3 && c->prm.track_residual) return ensure_partials(c);
    return LBM_OK;
}

}  // extern "C"
