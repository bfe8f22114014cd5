// lbm2d_host.inl -- host side of the D2Q9 path (dim = 2).  Included by lbm_b200.cu after lbm_ctx.h.
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
    if (c->slab_emulation) return LBM_OK;
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
    for (int a = 0; a < 4; a++)
        if (!c->out2[a] && (a < 3 || vmag)) { if ((rc = dev_alloc(c, &c->out2[a], c->FS))) return rc; }
    const int bx = block2d(c);
    const unsigned gx = (unsigned)((c->N + bx - 1) / bx);
    lbm2d::KP2 p = make_kp2(c, c->lat[c->cur], nullptr);
    p.u = c->out2[0]; p.v = c->out2[1]; p.rho = c->out2[2];
    p.vmag = vmag ? c->out2[3] : nullptr;
    lbm2d::k2_macro<false><<<dim3(gx, c->nyl), bx, 0, c->s_main>>>(p);
    c->launches++;
    CK(cudaGetLastError());
    const long long off = c->host_local ? 0 : (long long)c->j0 * c->N;
    double *host[4] = {u, v, rho, vmag};
    for (int a = 0; a < 4; a++)
        if (host[a])
            CK(cudaMemcpyAsync(host[a] + off, c->out2[a], (size_t)c->FS * 8, cudaMemcpyDeviceToHost, c->s_main));
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

