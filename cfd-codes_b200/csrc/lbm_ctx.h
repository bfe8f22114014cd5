// lbm_ctx.h -- context, error plumbing, NCCL loader and small shared helpers of the host side.
// Included once by lbm_b200.cu (single translation unit).
#pragma once
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

static std::mutex g_nccl_mutex;  // host programs create one context per thread concurrently

static int nccl_load() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
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
    cudaStream_t s_main = nullptr, s_halo = nullptr, s_copy = nullptr;   // s_copy: host<->device transfers (created on first use)
    bool own_main = false;
    cudaEvent_t ev_bnd = nullptr, ev_halo = nullptr, ev_main = nullptr, ev_int = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_up = nullptr, ev_mv = nullptr, ev_down = nullptr, ev_commit = nullptr;
    double *Fstage = nullptr, *side_stage = nullptr;   // upload staging (lbm_set_f_async): [19][FS] and the f_prev column [nzl][N]
    bool stage_ready = false, stage_has_prev = false, stage_commit_pending = false, down_pending = false;
    bool halo_pending = false;  // ev_halo has been recorded and not yet waited on by s_main

    int N = 0, nzl = 0, k0 = 0, k1 = 0;
    int G = 1;                            // ghost planes on each side of the slab in lat[]
    long long N2 = 0, PS = 0, FS = 0;
    double NU, TAU, OMEGA, OMEGAm, OmOMEGA, gammainv, ulid, vlid, wlid;

    double *lat_alloc[2] = {nullptr, nullptr};
    long long pad = 0;                    // elements of padding before/after each lattice
    double *lat[2] = {nullptr, nullptr};  // post-collision lattices (= lat_alloc + pad), ghost plane first
    int cur = 0;                          // lat[cur] holds S of the latest iteration
    double *Fbuf = nullptr;               // the reference's f2 of the latest iteration when f_valid
    bool f_valid = false;
    double *side = nullptr;               // [SIDE_DEPTH][nzl][N] stale-f14 history of the x = N-1 face
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
    double *out2[4] = {nullptr, nullptr, nullptr, nullptr};   // output staging: u, v, rho, vmag
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
    // graph launch.  The kernel arguments of iteration t depend on (lattice parity, t mod 4);
    // GRAPH_LEN is a multiple of 4, so a graph captured in one phase leaves the context in the same
    // phase and can be replayed back to back.
    int graphs = -1;                         // -1 auto (small grids), 0 off, 1 on
    cudaGraphExec_t graph_exec[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // two iterations per pass (k_step_pair): -1 auto, 0 off, 1 on; tile variant (PAIR_VARIANTS, -1 = default); planes per z segment
    int pair = -1, pair_variant = -1, pair_lz = 0;
    CUtensorMap tmap[2];                  // TMA-staged variants: one 4-D map per lattice (x, y, z, q)
    bool tmap_ok = false;
    int tmap_ty = 0;                      // box rows the maps were encoded for
    // device-side convergence loop (lbm3d_converge.inl): one conditional-WHILE graph, its parameters, the mapped host log
    cudaGraph_t conv_graph = nullptr;
    cudaGraphExec_t conv_exec = nullptr;
    long long conv_skip = 0, launches_per_block = 0;
    int conv_phase = -1, conv_mode = 1;
    struct ConvParams *conv_params = nullptr;
    struct ConvLog *conv_log = nullptr, *conv_log_dev = nullptr;
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
    long long halo_timeout_ms = 600000;   // direct halo: wall-clock limit of one wait for a neighbour (option halo_timeout_ms)
    bool host_local = false;  // host arrays are compact local slabs instead of full LU4/LU3 arrays
    ncclComm_t comm = nullptr;
    bool slab_emulation = false;   // LBM_B200_SLAB_EMULATION=1: one rank of a decomposition alone, no exchanges (ncu aid)
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

static int poll_flag(lbm_ctx *c) {
    int flag = 0;
    // the slab-boundary kernel of the last iteration runs on s_halo and may still raise a bit:
    // drain it BEFORE the flag is read
    if (c->s_halo) CK(cudaStreamSynchronize(c->s_halo));
    c->halo_pending = false;
    CK(cudaMemcpyAsync(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    if (flag & 4) return fail(LBM_ERR_CUDA, "direct halo: timed out waiting for a neighbour GPU's arrival flag");
    if (flag & 1) c->unstable = true;
    if (c->unstable) return fail(LBM_ERR_UNSTABLE, "Error: negative feq.  Therefore, unstable.");
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

