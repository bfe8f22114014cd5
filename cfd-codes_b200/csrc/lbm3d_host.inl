// lbm3d_host.inl -- host side of the D3Q19 path: launch geometry, kernel-parameter set-up, the
// per-iteration stream/event schedule (single GPU, NCCL halo, fused direct halo), CUDA-graph replay,
// materialisation of the reference's f2.  Included by lbm_b200.cu after lbm_ctx.h.
// Stale-f14 history of the x = N-1 face: SIDE_DEPTH slots of [nzl][N]; iteration t writes slot
// t % SIDE_DEPTH and reads the slot iteration t-2 wrote.  Depth 4 (not 3) so that the two fused
// iterations of k_step_pair read {t+2, t+3} and write {t, t+1}: disjoint (lbm3d_pair.cuh).
static const int SIDE_DEPTH = 4;
// Each slot has one ghost row below and above the slab ([nzl+2][N]): on several GPUs k_step_pair
// recomputes the first iteration on the neighbours' boundary planes and needs their history there.
static long long side_slot_elems(const lbm_ctx *c) { return (long long)(c->nzl + 2) * c->N; }
static double *side_slot(const lbm_ctx *c, long long t) {   // -> row of local plane 0
    return c->side + (((t % SIDE_DEPTH) + SIDE_DEPTH) % SIDE_DEPTH) * side_slot_elems(c) + c->N;
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
    p.G = c->G; p.zlo = 0; p.zhi = c->nzl; p.Lz = c->nzl; p.nseg1 = 1 << 30; p.zlo2 = 0; p.zhi2 = 0;
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

static int halo_exchange(lbm_ctx *c, double *L, cudaStream_t st, bool deep, long long next_t);

// ---- two iterations per pass (lbm3d_pair.cuh) -------------------------------------------------------
// `prepare` only raises the kernel's dynamic shared memory limit (done outside stream capture).
template <int TY, int MINB, int STAGE>
static int launch_pair_v(lbm_ctx *c, const lbm3d::KP &p, dim3 grid, cudaStream_t st, bool prepare) {
    const bool g1 = (c->prm.GAMMA == 1.0);
    const dim3 blk(lbm3d::PAIR_TX, TY, 1);
    const int smem = lbm3d::pair_smem_bytes(TY, STAGE);
#define LBM_PAIR_LAUNCH(M_, G_)                                                                              \
    do {                                                                                                     \
        static bool attr_done[16] = {false};                                                                 \
        if (!attr_done[c->dev & 15]) {                                                                       \
            CK(cudaFuncSetAttribute(lbm3d::k_step_pair<M_, G_, TY, MINB, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            attr_done[c->dev & 15] = true;                                                                   \
        }                                                                                                    \
        if (prepare) return LBM_OK;                                                                          \
        lbm3d::k_step_pair<M_, G_, TY, MINB, STAGE><<<grid, blk, smem, st>>>(p, c->tmap[c->cur]);                              \
    } while (0)
    if (c->prm.METHOD == 1) {
        if (g1) LBM_PAIR_LAUNCH(1, true); else LBM_PAIR_LAUNCH(1, false);
    } else {
        if (g1) LBM_PAIR_LAUNCH(0, true); else LBM_PAIR_LAUNCH(0, false);
    }
#undef LBM_PAIR_LAUNCH
    c->launches++;
    return LBM_OK;
}

// Tile variants: rows per block, resident blocks per SM, staging mode (lbm3d_pair.cuh).
struct PairVariant { int ty, minb, stage; };
static const PairVariant PAIR_VARIANTS[] = {{20, 1, 0}, {16, 1, 2}, {10, 2, 0}, {16, 1, 0}, {16, 1, 1}, {17, 1, 1}, {12, 1, 1}, {10, 2, 3}, {16, 1, 4}, {8, 2, 4}, {12, 1, 4}};
static const int PAIR_NVARIANTS = (int)(sizeof(PAIR_VARIANTS) / sizeof(PAIR_VARIANTS[0]));
// Defaults (profiles/r02_summary.md): 7 = two blocks of 32 x 10 per SM with direct pulls -- no restriction
// on N, best at 256^3 and on slabs; 8 = the TMA-staged 32 x 16 block with 64-plane segments for one big
// even-N grid per GPU (>= 2^26 cells: enough blocks to hide its one-block-per-SM wave quantisation), where
// it measured 1 % faster with 11 % fewer instructions.
static const int PAIR_VARIANT_DIRECT = 7, PAIR_VARIANT_TMA = 8;

static int pair_variant_requested(const lbm_ctx *c) {
    if (c->pair_variant >= 0 && c->pair_variant < PAIR_NVARIANTS) return c->pair_variant;
    return (c->prm.nranks == 1 && (c->N % 2) == 0 && c->N2 * c->N >= (1LL << 26)) ? PAIR_VARIANT_TMA : PAIR_VARIANT_DIRECT;
}

static int pair_variant(const lbm_ctx *c) {
    int v = pair_variant_requested(c);
    if (PAIR_VARIANTS[v].stage == 4 && !c->tmap_ok) v = PAIR_VARIANT_DIRECT;   // TMA needs 16-byte strides (N even)
    return v;
}

// One 4-D tensor map per lattice for the TMA-staged variants: dims (x, y, z incl. ghost planes, q),
// box = one 32 x TY tile of one population plane.  The encoder lives in libcuda; it is fetched through
// the runtime (cudaGetDriverEntryPoint), so the library needs no -lcuda.
typedef CUresult (*lbm_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tensor_maps(lbm_ctx *c) {
    c->tmap_ok = false;
    if (c->N % 2) return LBM_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return LBM_OK;   // no TMA variants; everything else works
    }
    const int ty = PAIR_VARIANTS[pair_variant_requested(c)].ty;
    for (int l = 0; l < 2; l++) {
        const cuuint64_t dims[4] = {(cuuint64_t)c->N, (cuuint64_t)c->N, (cuuint64_t)(c->nzl + 2 * c->G), (cuuint64_t)lbm3d::Q};
        const cuuint64_t strides[3] = {(cuuint64_t)c->N * 8, (cuuint64_t)c->N2 * 8, (cuuint64_t)c->PS * 8};
        const cuuint32_t box[4] = {(cuuint32_t)lbm3d::PAIR_BOX_X, (cuuint32_t)ty, 1, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = ((lbm_tmap_encode_fn)fn)(&c->tmap[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, c->lat[l], dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return LBM_OK;
    }
    c->tmap_ok = true;
    c->tmap_ty = ty;
    return LBM_OK;
}
static int pair_ty(const lbm_ctx *c) { return PAIR_VARIANTS[pair_variant(c)].ty; }

static int launch_pair(lbm_ctx *c, const lbm3d::KP &p, dim3 grid, cudaStream_t st, bool prepare = false) {
    switch (pair_variant(c)) {
        case 0: return launch_pair_v<20, 1, 0>(c, p, grid, st, prepare);
        case 1: return launch_pair_v<16, 1, 2>(c, p, grid, st, prepare);
        case 2: return launch_pair_v<10, 2, 0>(c, p, grid, st, prepare);
        case 3: return launch_pair_v<16, 1, 0>(c, p, grid, st, prepare);
        case 4: return launch_pair_v<16, 1, 1>(c, p, grid, st, prepare);
        case 5: return launch_pair_v<17, 1, 1>(c, p, grid, st, prepare);
        case 6: return launch_pair_v<12, 1, 1>(c, p, grid, st, prepare);
        case 7: return launch_pair_v<10, 2, 3>(c, p, grid, st, prepare);
        case 8: return launch_pair_v<16, 1, 4>(c, p, grid, st, prepare);
        case 9: return launch_pair_v<8, 2, 4>(c, p, grid, st, prepare);
        default: return launch_pair_v<12, 1, 4>(c, p, grid, st, prepare);
    }
}

// Every rank must come to the same answer (the halo depth depends on it), so the test uses only
// rank-independent quantities: the smallest slab of the decomposition and the average slab size.
static bool pair_enabled(const lbm_ctx *c) {
    if (c->prm.dim != 3) return false;
    const int P = c->prm.nranks;
    if (P > 1 && (c->p2p || c->N / P < 4 || c->G < 2)) return false;   // the fused direct halo is single-step only
    if (c->pair >= 0) return c->pair != 0;
    // automatic: HBM-resident slabs (>= 128^3 cells per GPU on average); below that the step is
    // launch-bound and the CUDA-graph / direct-halo paths of the single-step kernel win
    return (c->N2 * c->N) / P >= (1LL << 21);
}

// Iterations t = steps and t+1 in ONE pass: reads lat[cur] = S(t-1), writes lat[cur^1] = S(t+1).
static int run_pair(lbm_ctx *c) {
    lbm3d::KP p = make_kp(c);
    const long long t = c->steps;
    const int dsti = c->cur ^ 1;
    kp_set_lattices(c, &p, c->lat[c->cur], c->lat[dsti]);
    p.side_rd = side_slot(c, t + 2);  p.side_wr = side_slot(c, t);
    p.side_rd2 = side_slot(c, t + 3); p.side_wr2 = side_slot(c, t + 1);
    if (PAIR_VARIANTS[pair_variant(c)].stage == 4 && c->tmap_ty != pair_ty(c)) make_tensor_maps(c);
    const int ty = pair_ty(c);
    const unsigned gx = (c->N + lbm3d::PAIR_OX - 1) / lbm3d::PAIR_OX, gy = (c->N + ty - 3) / (ty - 2);
    int Lz = c->pair_lz > 0 ? c->pair_lz : (pair_variant(c) == PAIR_VARIANT_TMA ? 64 : 32);
    int rc;
    if (c->prm.nranks == 1) {
        p.zlo = 0; p.zhi = c->nzl;
        p.Lz = Lz < c->nzl ? Lz : c->nzl;
        if ((rc = launch_pair(c, p, dim3(gx, gy, (c->nzl + p.Lz - 1) / p.Lz), c->s_main))) return rc;
    } else {
        // Same two-stream schedule as run_iteration: s_halo (high priority) computes the two output
        // planes at each end of the slab -- everything the neighbours need -- and then exchanges
        // them; s_main computes the planes in between, concurrently.
        //   interior(s) reads planes [0, nzl) and writes [2, nzl-2): needs boundary(s-1);
        //   boundary(s) reads planes [-2, 4) + [nzl-4, nzl+2) and writes [0,2) + [nzl-2,nzl): needs
        //               interior(s-1) and the exchange of step s-1 (stream order on s_halo).
        CK(cudaStreamWaitEvent(c->s_main, c->ev_bnd, 0));
        CK(cudaEventRecord(c->ev_main, c->s_main));
        CK(cudaStreamWaitEvent(c->s_halo, c->ev_main, 0));
        p.Lz = 2; p.zlo = 0; p.zhi = 2; p.nseg1 = 1; p.zlo2 = c->nzl - 2; p.zhi2 = c->nzl;
        if ((rc = launch_pair(c, p, dim3(gx, gy, 2), c->s_halo))) return rc;
        CK(cudaEventRecord(c->ev_bnd, c->s_halo));
        if ((rc = halo_exchange(c, c->lat[dsti], c->s_halo, true, t + 2))) return rc;
        CK(cudaEventRecord(c->ev_halo, c->s_halo));
        c->halo_pending = true;
        const int zi = c->nzl - 4;   // interior output planes [2, nzl-2)
        if (zi > 0) {
            p.zlo = 2; p.zhi = c->nzl - 2; p.nseg1 = 1 << 30;
            const int nseg = (zi + Lz - 1) / Lz;
            p.Lz = (zi + nseg - 1) / nseg;   // equal segments
            if ((rc = launch_pair(c, p, dim3(gx, gy, nseg), c->s_main))) return rc;
        }
    }
    CK(cudaGetLastError());
    c->cur ^= 1;
    c->steps += 2;
    c->f_valid = false;
    c->residual_ready = false;
    return LBM_OK;
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
//
// `deep` (two-iterations-per-pass stepping): the next kernel recomputes the first of its two
// iterations on the neighbours' boundary planes, so two ghost planes per side are filled -- the
// outer one with the 5 populations that enter the inner one, the inner one with those 5 plus the
// 9 in-plane populations (19 planes per face and direction instead of 5, once per TWO iterations)
// -- and the x = N-1 stale-f14 row of the slot that kernel reads.
static const int Q_IN[9] = {0, 1, 2, 3, 4, 7, 8, 13, 14};   // cz = 0
static int halo_exchange(lbm_ctx *c, double *L, cudaStream_t st, bool deep, long long next_t) {
    if (c->slab_emulation) return LBM_OK;
    const int r = c->prm.rank, P = c->prm.nranks, G = c->G;
    const size_t n = (size_t)c->N2;
    const long long PS = c->PS, N2 = c->N2, nzl = c->nzl;
#define PLANE(q, kl) (L + (long long)(q) * PS + ((long long)(kl) + G) * N2)
    NK(g_nccl.GroupStart());
    if (!deep) {
        for (int a = 0; a < 5; a++) {
            if (r + 1 < P) {
                NK(g_nccl.Send(PLANE(Q_UP[a], nzl - 1), n, ncclDouble, r + 1, c->comm, st));
                NK(g_nccl.Recv(PLANE(Q_DN[a], nzl), n, ncclDouble, r + 1, c->comm, st));
            }
            if (r > 0) {
                NK(g_nccl.Send(PLANE(Q_DN[a], 0), n, ncclDouble, r - 1, c->comm, st));
                NK(g_nccl.Recv(PLANE(Q_UP[a], -1), n, ncclDouble, r - 1, c->comm, st));
            }
        }
    } else {
        double *side = side_slot(c, next_t - 2);   // the slot the first iteration of a pair starting at next_t reads
        if (r + 1 < P) {   // my top two planes -> rank+1's bottom ghosts; its bottom two planes -> my top ghosts
            for (int a = 0; a < 5; a++) NK(g_nccl.Send(PLANE(Q_UP[a], nzl - 2), n, ncclDouble, r + 1, c->comm, st));
            for (int a = 0; a < 5; a++) NK(g_nccl.Send(PLANE(Q_UP[a], nzl - 1), n, ncclDouble, r + 1, c->comm, st));
            for (int a = 0; a < 9; a++) NK(g_nccl.Send(PLANE(Q_IN[a], nzl - 1), n, ncclDouble, r + 1, c->comm, st));
            NK(g_nccl.Send(side + (nzl - 1) * c->N, (size_t)c->N, ncclDouble, r + 1, c->comm, st));
            for (int a = 0; a < 5; a++) NK(g_nccl.Recv(PLANE(Q_DN[a], nzl + 1), n, ncclDouble, r + 1, c->comm, st));
            for (int a = 0; a < 5; a++) NK(g_nccl.Recv(PLANE(Q_DN[a], nzl), n, ncclDouble, r + 1, c->comm, st));
            for (int a = 0; a < 9; a++) NK(g_nccl.Recv(PLANE(Q_IN[a], nzl), n, ncclDouble, r + 1, c->comm, st));
            NK(g_nccl.Recv(side + nzl * c->N, (size_t)c->N, ncclDouble, r + 1, c->comm, st));
        }
        if (r > 0) {
            for (int a = 0; a < 5; a++) NK(g_nccl.Send(PLANE(Q_DN[a], 1), n, ncclDouble, r - 1, c->comm, st));
            for (int a = 0; a < 5; a++) NK(g_nccl.Send(PLANE(Q_DN[a], 0), n, ncclDouble, r - 1, c->comm, st));
            for (int a = 0; a < 9; a++) NK(g_nccl.Send(PLANE(Q_IN[a], 0), n, ncclDouble, r - 1, c->comm, st));
            NK(g_nccl.Send(side, (size_t)c->N, ncclDouble, r - 1, c->comm, st));
            for (int a = 0; a < 5; a++) NK(g_nccl.Recv(PLANE(Q_UP[a], -2), n, ncclDouble, r - 1, c->comm, st));
            for (int a = 0; a < 5; a++) NK(g_nccl.Recv(PLANE(Q_UP[a], -1), n, ncclDouble, r - 1, c->comm, st));
            for (int a = 0; a < 9; a++) NK(g_nccl.Recv(PLANE(Q_IN[a], -1), n, ncclDouble, r - 1, c->comm, st));
            NK(g_nccl.Recv(side - c->N, (size_t)c->N, ncclDouble, r - 1, c->comm, st));
        }
    }
    NK(g_nccl.GroupEnd());
#undef PLANE
    return LBM_OK;
}

static const int GRAPH_LEN = 60;

static void drop_graphs(lbm_ctx *c) {
    for (int k = 0; k < 8; k++)
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
    const int phase = (int)(c->steps % SIDE_DEPTH) * 2 + c->cur;
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
    c->steps += GRAPH_LEN;        // GRAPH_LEN is a multiple of 4: cur and the side slot phase are unchanged
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
    p.side_rd = side_slot(c, t + 2);  // written by iteration t-2
    p.side_wr = side_slot(c, t);
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
        const bool deep = pair_enabled(c);   // the next step may be a fused pair: fill both ghost planes (halo_exchange)
        if (!c->p2p) {
            launch_step(c, mode, p, dim3(gxy.x, gxy.y, nb), blk, c->s_halo);
            CK(cudaEventRecord(c->ev_bnd, c->s_halo));
            if (!deep) {
                int rc = halo_exchange(c, c->lat[src_i ^ 1], c->s_halo, false, t + 1);
                if (rc) return rc;
            }
        } else {
            // Direct NVLink halo: wait until both neighbours finished their boundary kernels of the
            // previous iteration (they filled my ghosts and are done reading the ghosts I overwrite
            // now), then ONE kernel computes the boundary planes, stores the face-crossing
            // populations into the neighbours' ghost planes and publishes this iteration's stamp.
            const long long rel = t - c->p2p_base;
            const int r = c->prm.rank, P = c->prm.nranks;
            lbm3d::k_wait_arrivals<<<1, 1, 0, c->s_halo>>>(c->sig, r > 0 ? rel : 0, r + 1 < P ? rel : 0, c->d_flag,
                                                            (unsigned long long)c->halo_timeout_ms * 1000000ULL);
            c->launches++;
            const int dsti = src_i ^ 1;
            for (int a = 0; a < 5; a++) {
                p.peer_dn[a] = r > 0 ? c->peer_lat[0][dsti] + (long long)Q_DN[a] * c->peer_PS[0] + (long long)(c->peer_nzl[0] + c->G) * c->N2
                                     : nullptr;
                p.peer_up[a] = r + 1 < P ? c->peer_lat[1][dsti] + (long long)Q_UP[a] * c->peer_PS[1] + (long long)(c->G - 1) * c->N2 : nullptr;
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
        if (c->nzl > 2) {
            p.kl0 = 1; p.kstride = 1; p.partial_base = (int)(bpp * nb);
            launch_step(c, mode, p, dim3(gxy.x, gxy.y, c->nzl - 2), blk, c->s_main);
        }
        if (deep && !c->p2p) {
            // the deep exchange also ships planes 1 and nzl-2, which the interior launch computed:
            // no overlap for the (rare) single steps between fused pairs
            CK(cudaEventRecord(c->ev_int, c->s_main));
            CK(cudaStreamWaitEvent(c->s_halo, c->ev_int, 0));
            int rc = halo_exchange(c, c->lat[src_i ^ 1], c->s_halo, true, t + 1);
            if (rc) return rc;
        }
        CK(cudaEventRecord(c->ev_halo, c->s_halo));
        c->halo_pending = true;
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

