// lbm3d_converge.inl -- the reference's convergence loop (3D:103-124) WITHOUT host synchronisation.
//
//   for (t = 1; t < MAXITER; t++) { macrocollideandstream(); HechtBC(f2);
//       if (t % skip == 0) { df = diff(f2, f1, NQ) * df0; print; if (df < THRESH) break; }  swap; }
//
// The whole loop is ONE CUDA graph: a conditional WHILE node whose body is a residual block -- `skip`
// iterations (plain / fused-pair launches, the STOREF and DIFF_STOREF iterations, the partial-sum
// kernel) followed by k_converge_check, which evaluates `df = sqrt(sum) * df0`, appends the block's
// record to a ring in MAPPED HOST MEMORY and sets the loop condition on the device
// (cudaGraphSetConditional).  The host launches the graph once and only polls that ring to report the
// blocks as they finish; it never calls a CUDA synchronisation function until the loop has ended.
// Included by lbm_b200.cu after lbm3d_host.inl.

static const int CONV_RING = 4096;   // records the host may lag behind by

struct ConvParams {       // device memory, refreshed before every launch (the graph itself is reusable)
    double df0, thresh;
    long long skip, max_blocks, t_first;   // t_first: iteration index of the first block's check minus skip
    long long blocks;                      // blocks done so far (device counter)
};

struct ConvLog {          // mapped, pinned host memory
    volatile long long seq;                // number of records published
    lbm_block_record rec[CONV_RING];
};

__global__ void k_converge_check(ConvParams *cp, const double *d_red, const int *flag, ConvLog *log,
                                 cudaGraphConditionalHandle handle) {
    const long long b = ++cp->blocks;
    const double df = sqrt(d_red[0]) * cp->df0;          // 3D:684, 112
    const bool unstable = (*flag & 1) != 0;
    const bool stop = (df < cp->thresh) || unstable || (b >= cp->max_blocks);   // 3D:119-122
    lbm_block_record r;
    r.iteration = cp->t_first + b * cp->skip;
    r.df = df;
    r.status = unstable ? LBM_ERR_UNSTABLE : LBM_OK;
    r.last = stop ? 1 : 0;
    log->rec[(b - 1) % CONV_RING] = r;
    __threadfence_system();
    log->seq = b;
    __threadfence_system();
    cudaGraphSetConditional(handle, stop ? 0u : 1u);
}

static void drop_converge_graph(lbm_ctx *c) {
    if (c->conv_exec) { cudaGraphExecDestroy(c->conv_exec); c->conv_exec = nullptr; }
    if (c->conv_graph) { cudaGraphDestroy(c->conv_graph); c->conv_graph = nullptr; }
}

static bool converge_on_device(const lbm_ctx *c, long long skip) {
    // body replays must leave (lattice parity, stale-f14 slot phase) unchanged: skip % 4 == 0
    return c->prm.dim == 3 && c->prm.nranks == 1 && c->prm.track_residual && skip >= 4 && (skip % SIDE_DEPTH) == 0 &&
           c->conv_mode != 0;
}

// One residual block as stream work (capturable): skip-2 plain iterations, STOREF, DIFF_STOREF.
// The body is replayed, so it must leave the lattice parity as it found it: every launch (a fused pair
// or a single iteration) flips `cur` once, and the number of launches is made even by running the
// last two plain iterations singly when needed.
static int enqueue_block(lbm_ctx *c, long long skip) {
    int rc;
    const bool use_pairs = pair_enabled(c);
    long long n_pairable = skip - 2;
    if (use_pairs && ((n_pairable / 2 + n_pairable % 2 + 2) % 2) != 0) n_pairable -= 2;
    for (long long t = 0; t < skip; t++) {
        int mode = lbm3d::MODE_STEP;
        if (t == skip - 1) mode = lbm3d::MODE_STEP_DIFF_STOREF;
        else if (t == skip - 2) mode = lbm3d::MODE_STEP_STOREF;
        if (use_pairs && t + 2 <= n_pairable) {
            if ((rc = run_pair(c))) return rc;
            t += 1;
        } else if ((rc = run_iteration(c, mode))) {
            return rc;
        }
    }
    return LBM_OK;
}

static int build_converge_graph(lbm_ctx *c, long long skip) {
    drop_converge_graph(c);
    int rc;
    if (!c->conv_params) {
        if ((rc = dev_alloc(c, &c->conv_params, 1))) return rc;
        CK(cudaHostAlloc((void **)&c->conv_log, sizeof(ConvLog), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void **)&c->conv_log_dev, c->conv_log, 0));
    }
    if (pair_enabled(c)) {   // raise the pair kernel's shared-memory limit outside the capture
        lbm3d::KP p = make_kp(c);
        if ((rc = launch_pair(c, p, dim3(1, 1, 1), c->s_main, true))) return rc;
    }
    CK(cudaGraphCreate(&c->conv_graph, 0));
    cudaGraphConditionalHandle handle;
    CK(cudaGraphConditionalHandleCreate(&handle, c->conv_graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, c->conv_graph, nullptr, 0, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];

    const long long steps0 = c->steps, launches0 = c->launches;
    const int cur0 = c->cur;
    const bool fvalid0 = c->f_valid, rr0 = c->residual_ready;
    CK(cudaStreamBeginCaptureToGraph(c->s_main, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_block(c, skip);
    if (!rc) {
        k_converge_check<<<1, 1, 0, c->s_main>>>(c->conv_params, c->d_red, c->d_flag, c->conv_log_dev, handle);
        if (cudaGetLastError() != cudaSuccess) rc = fail(LBM_ERR_CUDA, "k_converge_check launch failed");
    }
    cudaGraph_t captured = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->s_main, &captured);
    c->launches_per_block = c->launches - launches0 + 1;
    c->steps = steps0; c->launches = launches0; c->cur = cur0; c->f_valid = fvalid0; c->residual_ready = rr0;   // nothing ran
    if (rc) { drop_converge_graph(c); return rc; }
    if (e != cudaSuccess) { drop_converge_graph(c); return fail(LBM_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e)); }
    e = cudaGraphInstantiate(&c->conv_exec, c->conv_graph, 0);
    if (e != cudaSuccess) { drop_converge_graph(c); return fail(LBM_ERR_CUDA, "cudaGraphInstantiate (while loop): %s", cudaGetErrorString(e)); }
    c->conv_skip = skip;
    c->conv_phase = (int)(steps0 % SIDE_DEPTH) * 2 + cur0;
    return LBM_OK;
}

// Host fallback (several GPUs, skip not a multiple of 4, option converge_on_device = 0): the same loop with
// one lbm_residual synchronisation per block.
static int converge_host_loop(lbm_ctx *c, long long skip, double thresh, long long max_blocks, double df0,
                              lbm_block_fn on_block, void *user, long long *blocks_done, double *df_last) {
    long long b = 0;
    int rc = LBM_OK;
    double df = 0.0;
    while (b < max_blocks) {
        if ((rc = lbm_step(c, skip))) break;
        double d;
        rc = lbm_residual(c, &d);
        if (rc && rc != LBM_ERR_UNSTABLE) break;
        df = d * df0;
        b++;
        lbm_block_record r;
        r.iteration = c->steps - 1;
        r.df = df;
        r.status = rc;
        r.last = (rc || df < thresh || b >= max_blocks) ? 1 : 0;
        if (on_block) on_block(&r, user);
        if (r.last) break;
    }
    if (blocks_done) *blocks_done = b;
    if (df_last) *df_last = df;
    return rc;
}
