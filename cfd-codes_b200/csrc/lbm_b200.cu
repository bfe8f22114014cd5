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

#include <mutex>
#include <vector>

#include "../../include/lbm_b200.h"
#include "lbm3d_kernels.cuh"
#include "lbm3d_pair.cuh"
#include "lbm2d_kernels.cuh"
#include "lbm_ctx.h"
#include "lbm3d_host.inl"
#include "lbm2d_host.inl"
#include "lbm3d_converge.inl"

// ---- C ABI ---------------------------------------------------------------------------------------
extern "C" {

const char *lbm_last_error(void) { return g_err; }

const char *lbm_version(void) { return "lbm_b200 0.2 (sm_100a, fp64, -fmad=false, D3Q19 pull/two-lattice, two iterations per pass)"; }

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
    bool peers_quiet = true;
    if (c->p2p && c->s_halo && c->steps > c->p2p_base) {
        // the neighbours' last boundary kernels store into my ghost planes: let them land first
        const long long rel = c->steps - c->p2p_base;
        lbm3d::k_wait_arrivals<<<1, 1, 0, c->s_halo>>>(c->sig, c->prm.rank > 0 ? rel : 0,
                                                        c->prm.rank + 1 < c->prm.nranks ? rel : 0, c->d_flag,
                                                        (unsigned long long)c->halo_timeout_ms * 1000000ULL);
        int flag = 0;
        if (cudaStreamSynchronize(c->s_halo) == cudaSuccess &&
            cudaMemcpy(&flag, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && (flag & 4))
            peers_quiet = false;
    }
    if (c->s_main) cudaStreamSynchronize(c->s_main);
    if (c->s_halo) cudaStreamSynchronize(c->s_halo);
    if (!peers_quiet) {
        // A neighbour never reported its last boundary kernel: it may still be storing into this
        // context's ghost planes.  Freeing them now could hand that memory to somebody else, so the
        // device buffers are deliberately leaked and the caller is told.
        if (c->comm) g_nccl.CommDestroy(c->comm);
        delete c;
        return fail(LBM_ERR_CUDA, "lbm_destroy: a neighbour GPU did not finish within halo_timeout_ms; device memory of this context was not freed");
    }
    for (int side = 0; side < 2; side++)
        for (int a = 0; a < 3; a++)
            if (c->peer_ipc[side][a]) cudaIpcCloseMemHandle(c->peer_ipc[side][a]);
    cudaFree(c->sig); cudaFree(c->done_count);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->lat_alloc[0]); cudaFree(c->lat_alloc[1]); cudaFree(c->Fbuf); cudaFree(c->side);
    cudaFree(c->partials); cudaFree(c->d_red); cudaFree(c->d_flag);
    for (int a = 0; a < 5; a++) cudaFree(c->d_out[a]);
    cudaFree(c->Fstage); cudaFree(c->side_stage);
    if (c->ev_up) cudaEventDestroy(c->ev_up);
    if (c->ev_mv) cudaEventDestroy(c->ev_mv);
    if (c->ev_down) cudaEventDestroy(c->ev_down);
    if (c->ev_commit) cudaEventDestroy(c->ev_commit);
    if (c->s_copy) { cudaStreamSynchronize(c->s_copy); cudaStreamDestroy(c->s_copy); }
    cudaFree(c->u2); cudaFree(c->v2); cudaFree(c->rho2);
    for (int a = 0; a < 4; a++) cudaFree(c->out2[a]);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_int) cudaEventDestroy(c->ev_int);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    drop_graphs(c);
    drop_converge_graph(c);
    cudaFree(c->conv_params);
    if (c->conv_log) cudaFreeHost(c->conv_log);
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
    c->G = p.nranks > 1 ? 2 : 1;   // two ghost planes per side on several GPUs (k_step_pair recomputes one plane of iteration t)
    c->PS = (long long)(c->nzl + 2 * c->G) * c->N2;
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
    const long long side_n = side_slot_elems(c);
    if ((rc = dev_alloc(c, &c->side, SIDE_DEPTH * side_n))) return rc;
    if ((rc = dev_alloc(c, &c->d_red, 2))) return rc;
    if ((rc = dev_alloc(c, &c->d_flag, 1))) return rc;
    CK(cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->s_main));
    CK(cudaMemsetAsync(c->d_red, 0, 2 * sizeof(double), c->s_main));
    CK(cudaMallocHost((void **)&c->h_pin, 4 * sizeof(double)));

    // stale-f14 history starts at the initial value w[14] = 1/36 (3D:192-205)
    lbm3d::k_fill<<<(unsigned)((SIDE_DEPTH * side_n + 255) / 256), 256, 0, c->s_main>>>(c->side, SIDE_DEPTH * side_n, LBM_W2);
    c->launches++;

    // S(-1) = collide(f = w) in both lattices incl. ghost planes (colliding the initial equilibrium
    // is NOT a no-op in fp64: sum of w[q] != 1)
    dim3 blk, gxy;
    choose_block(c, &blk, &gxy);
    for (int l = 0; l < 2; l++) {
        lbm3d::KP kp = make_kp(c);
        kp_set_lattices(c, &kp, nullptr, c->lat[l]);
        kp.Fbuf = nullptr;
        dim3 grid(gxy.x, gxy.y, c->nzl + 2 * c->G);
        if (p.METHOD == 1) lbm3d::k_collide_cells<1><<<grid, blk, 0, c->s_main>>>(kp, 1);
        else               lbm3d::k_collide_cells<0><<<grid, blk, 0, c->s_main>>>(kp, 1);
        c->launches++;
    }
    CK(cudaGetLastError());
    c->cur = 0;
    c->steps = 0;
    if ((rc = make_tensor_maps(c))) return rc;
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
        CKD(cudaEventCreateWithFlags(&c->ev_int, cudaEventDisableTiming));
        // Profiling aid (tools/README.md): LBM_B200_SLAB_EMULATION=1 runs ONE rank of a multi-GPU
        // decomposition without its neighbours -- same slab, same launches, exchanges skipped (the ghost
        // planes keep their initial values, so the RESULTS ARE NOT THE CAVITY'S) -- so that ncu, which must
        // not wrap a multi-rank command, can capture the slab kernels on a single GPU.
        const char *emu = getenv("LBM_B200_SLAB_EMULATION");
        c->slab_emulation = emu && emu[0] == '1';
        if (!c->slab_emulation) {
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
    // The leading iterations run the plain step kernel: all but the last two when the residual is
    // tracked (STOREF, DIFF_STOREF), all but the last one otherwise -- the final iteration is always
    // a single step, so that lat[cur^1] holds S(t-1) and the reference's f2 can be re-materialised.
    const bool use_pairs = pair_enabled(c);
    const long long n_plain = track ? n - 2 : (use_pairs ? n - 1 : n);
    lbm_ctx::Span sp{nullptr, nullptr, 0};
    if (n_plain > 0) {
        if (c->spans.size() > 256) resolve_spans(c, false);
        sp.a = span_event(c);
        sp.b = span_event(c);
        if (!sp.a || !sp.b) return fail(LBM_ERR_CUDA, "cudaEventCreate failed");
        sp.iters = n_plain;
        CK(cudaEventRecord(sp.a, c->s_main));
    }
    const bool use_graphs = graphs_enabled(c) && !use_pairs;
    for (long long t = 0; t < n; t++) {
        int mode = lbm3d::MODE_STEP;
        if (track) {
            if (t == n - 1) mode = lbm3d::MODE_STEP_DIFF_STOREF;     // F(t) vs F(t-1): the reference's diff(f2, f1)
            else if (t == n - 2) mode = lbm3d::MODE_STEP_STOREF;     // leaves F(t-1) for the last iteration
        }
        if (use_graphs && t + GRAPH_LEN <= n_plain) {
            if ((rc = run_graph_block(c))) return rc;
            t += GRAPH_LEN - 1;
        } else if (use_pairs && t + 2 <= n_plain) {
            if ((rc = run_pair(c))) return rc;                       // two iterations, one pass over HBM
            t += 1;
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

int lbm_converge(lbm_ctx *c, long long skip, double thresh, long long max_blocks, double df0, lbm_block_fn on_block,
                 void *user, long long *blocks_done, double *df_last) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (skip < 1 || max_blocks < 0) return fail(LBM_ERR_INVALID, "lbm_converge: bad skip / max_blocks");
    if (blocks_done) *blocks_done = 0;
    if (df_last) *df_last = 0.0;
    if (max_blocks == 0) return LBM_OK;
    if (c->unstable) return fail(LBM_ERR_UNSTABLE, "Error: negative feq.  Therefore, unstable.");
    CK(cudaSetDevice(c->dev));
    if (!converge_on_device(c, skip)) return converge_host_loop(c, skip, thresh, max_blocks, df0, on_block, user, blocks_done, df_last);

    int rc;
    if ((rc = ensure_Fbuf(c))) return rc;
    if ((rc = ensure_partials(c))) return rc;
    const int phase = (int)(c->steps % SIDE_DEPTH) * 2 + c->cur;
    if (!c->conv_exec || c->conv_skip != skip || c->conv_phase != phase) {
        if ((rc = build_converge_graph(c, skip))) return rc;
    }
    ConvParams hp;
    hp.df0 = df0; hp.thresh = thresh; hp.skip = skip; hp.max_blocks = max_blocks;
    hp.t_first = c->steps - 1;   // the first block's check is iteration (steps - 1) + skip
    hp.blocks = 0;
    c->conv_log->seq = 0;
    CK(cudaMemcpyAsync(c->conv_params, &hp, sizeof hp, cudaMemcpyHostToDevice, c->s_main));
    CK(cudaGraphLaunch(c->conv_exec, c->s_main));
    // Report the blocks as their records arrive in mapped host memory: plain loads, no CUDA call.
    long long seen = 0;
    double df = 0.0;
    int status = LBM_OK;
    bool done = false;
    unsigned spins = 0;
    while (!done) {
        const long long seq = c->conv_log->seq;
        if (seq == seen) {
            if ((++spins & 0xfff) == 0 && cudaStreamQuery(c->s_main) != cudaErrorNotReady) {
                // the graph ended (or failed) without publishing a final record: leave through the sync below
                if (c->conv_log->seq == seen) break;
            }
            continue;
        }
        __sync_synchronize();
        for (; seen < seq; seen++) {
            const lbm_block_record r = c->conv_log->rec[seen % CONV_RING];
            df = r.df;
            status = r.status;
            if (on_block) on_block(&r, user);
            if (r.last) done = true;
        }
    }
    cudaGetLastError();
    CK(cudaStreamSynchronize(c->s_main));   // the loop has ended: one synchronisation, after convergence
    c->steps += seen * skip;
    c->launches += seen * c->launches_per_block;
    c->f_valid = true;           // the last iteration of a block stores the reference's f2
    c->residual_ready = true;
    if (blocks_done) *blocks_done = seen;
    if (df_last) *df_last = df;
    if (status == LBM_ERR_UNSTABLE) { c->unstable = true; return fail(LBM_ERR_UNSTABLE, "Error: negative feq.  Therefore, unstable."); }
    return poll_flag(c);
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
    if (c->prm.nranks > 1 && !c->slab_emulation) {
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

// ---- host <-> device state transfer ----------------------------------------------------------------
// Uploads go through a second device buffer (`Fstage`) on a dedicated copy stream, downloads of the
// macroscopic fields likewise, so that a host which processes a SEQUENCE of states can overlap the PCIe
// traffic of batch k+1 / k-1 with the iterations of batch k:
//     lbm_set_f_async(next)  ||  lbm_step(...), lbm_residual(...)   then   lbm_macrovars_async(out)
// lbm_set_f / lbm_macrovars are the same operations followed by a wait.
static int ensure_copy_stream(lbm_ctx *c) {
    if (c->s_copy) return LBM_OK;
    CK(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_mv, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_down, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_commit, cudaEventDisableTiming));
    return LBM_OK;
}

int lbm_set_f_async(lbm_ctx *c, const double *host_f, const double *host_f_prev) {
    if (!c || !host_f) return fail(LBM_ERR_INVALID, "null argument");
    if (c->prm.dim != 3) return fail(LBM_ERR_INVALID, "lbm_set_f_async: dim=3 only (use lbm_set_f)");
    CK(cudaSetDevice(c->dev));
    int rc = ensure_copy_stream(c);
    if (rc) return rc;
    if (!c->Fstage && (rc = dev_alloc(c, &c->Fstage, (long long)lbm3d::Q * c->FS))) return rc;
    if (!c->side_stage && (rc = dev_alloc(c, &c->side_stage, (long long)c->nzl * c->N))) return rc;
    if (c->stage_commit_pending) {   // the previous commit still reads Fstage's predecessor: order after it
        CK(cudaStreamWaitEvent(c->s_copy, c->ev_commit, 0));
        c->stage_commit_pending = false;
    }
    const long long N3 = c->host_local ? c->FS : c->N2 * c->N;
    const long long hk0 = c->host_local ? 0 : c->k0;
    for (int q = 0; q < lbm3d::Q; q++)
        CK(cudaMemcpyAsync(c->Fstage + (long long)q * c->FS, host_f + (long long)q * N3 + hk0 * c->N2,
                           (size_t)c->FS * sizeof(double), cudaMemcpyHostToDevice, c->s_copy));
    c->stage_has_prev = host_f_prev != nullptr;
    if (host_f_prev)   // only the x = N-1 column of population 14 is needed from the other buffer (3D:421)
        CK(cudaMemcpy2DAsync(c->side_stage, sizeof(double), host_f_prev + 14LL * N3 + hk0 * c->N2 + (c->N - 1),
                             (size_t)c->N * sizeof(double), sizeof(double), (size_t)c->nzl * c->N, cudaMemcpyHostToDevice,
                             c->s_copy));
    CK(cudaEventRecord(c->ev_up, c->s_copy));
    c->stage_ready = true;
    return LBM_OK;
}

int lbm_set_f_commit(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (!c->stage_ready) return fail(LBM_ERR_INVALID, "lbm_set_f_commit without lbm_set_f_async");
    CK(cudaSetDevice(c->dev));
    if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
    CK(cudaStreamWaitEvent(c->s_main, c->ev_up, 0));
    // the uploaded state becomes the F buffer; the old F buffer becomes the next staging area
    double *t = c->Fbuf; c->Fbuf = c->Fstage; c->Fstage = t;
    c->stage_ready = false;
    // Stale f14 of the x = N-1 face (3D:421): iteration t = steps reads the slot iteration t-2 wrote --
    // that value lives in the reference's OTHER buffer (f_prev) -- and iteration t+1 reads what t-1
    // wrote: this state's own post-BC f14, gathered on the device from the uploaded F buffer.
    const long long side_n = (long long)c->nzl * c->N;
    const unsigned gb = (unsigned)((side_n + 255) / 256);
    lbm3d::k_gather_xmax<<<gb, 256, 0, c->s_main>>>(c->Fbuf + 14LL * c->FS, c->N, side_n, side_slot(c, c->steps + 3));
    c->launches++;
    if (c->stage_has_prev) {
        CK(cudaMemcpyAsync(side_slot(c, c->steps + 2), c->side_stage, (size_t)side_n * sizeof(double), cudaMemcpyDeviceToDevice, c->s_main));
    } else {
        lbm3d::k_gather_xmax<<<gb, 256, 0, c->s_main>>>(c->Fbuf + 14LL * c->FS, c->N, side_n, side_slot(c, c->steps + 2));
        c->launches++;
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
        int rc = halo_exchange(c, c->lat[c->cur], c->s_main, pair_enabled(c), c->steps);
        if (rc) return rc;
    }
    if (c->Fstage) {   // the next upload overwrites the buffer that was Fbuf: everything queued so far may still read it
        CK(cudaEventRecord(c->ev_commit, c->s_main));
        c->stage_commit_pending = true;
    }
    c->f_valid = true;
    c->residual_ready = false;
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
    if ((rc = lbm_set_f_async(c, host_f, host_f_prev))) return rc;
    if ((rc = lbm_set_f_commit(c))) return rc;
    CK(cudaStreamSynchronize(c->s_main));
    return LBM_OK;
}

// k_macrovars on s_main, then the five field downloads on the copy stream.
int lbm_macrovars_async(lbm_ctx *c, double *u1, double *u2, double *u3, double *rho, double *vmag) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (c->prm.dim != 3) return fail(LBM_ERR_INVALID, "lbm_macrovars_async: dim=3 only (use lbm_macrovars)");
    CK(cudaSetDevice(c->dev));
    int rc = ensure_copy_stream(c);
    if (rc) return rc;
    if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
    if ((rc = materialise_F(c))) return rc;
    double *host[5] = {u1, u2, u3, rho, vmag};
    for (int a = 0; a < 5; a++)
        if (!c->d_out[a]) { if ((rc = dev_alloc(c, &c->d_out[a], c->FS))) return rc; }
    const long long n = c->FS;
    if (c->down_pending) CK(cudaStreamWaitEvent(c->s_main, c->ev_down, 0));   // previous downloads still read d_out
    lbm3d::k_macrovars<<<(unsigned)((n + 255) / 256), 256, 0, c->s_main>>>(c->Fbuf, c->FS, c->d_out[0], c->d_out[1],
                                                                          c->d_out[2], c->d_out[3], c->d_out[4]);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev_mv, c->s_main));
    CK(cudaStreamWaitEvent(c->s_copy, c->ev_mv, 0));
    for (int a = 0; a < 5; a++)
        if (host[a])
            CK(cudaMemcpyAsync(host[a] + (c->host_local ? 0LL : (long long)c->k0 * c->N2), c->d_out[a], (size_t)n * sizeof(double),
                               cudaMemcpyDeviceToHost, c->s_copy));
    CK(cudaEventRecord(c->ev_down, c->s_copy));
    c->down_pending = true;
    return LBM_OK;
}

int lbm_io_wait(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->dev));
    if (c->s_copy) CK(cudaStreamSynchronize(c->s_copy));
    c->down_pending = false;
    return LBM_OK;
}

int lbm_macrovars(lbm_ctx *c, double *u1, double *u2, double *u3, double *rho, double *vmag) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->dev));
    if (c->prm.dim == 2) {
        if (c->halo_pending) { CK(cudaStreamWaitEvent(c->s_main, c->ev_halo, 0)); c->halo_pending = false; }
        return macrovars2d(c, u1, u2, rho, vmag);
    }
    int rc = lbm_macrovars_async(c, u1, u2, u3, rho, vmag);
    if (rc) return rc;
    return lbm_io_wait(c);
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
    o->pair_enabled = (c->prm.dim == 3 && pair_enabled(c)) ? 1 : 0;
    o->pair_variant = c->prm.dim == 3 ? pair_variant(c) : -1;
    return LBM_OK;
}

int lbm_set_option(lbm_ctx *c, const char *key, long long value) {
    if (!c || !key) return fail(LBM_ERR_INVALID, "null argument");
    drop_graphs(c);  // captured launches bake in the block shape and kernel variant
    drop_converge_graph(c);
    if (!strcmp(key, "graphs")) {
        c->graphs = value < 0 ? -1 : (value != 0);
        return LBM_OK;
    }
    if (!strcmp(key, "block_x") || !strcmp(key, "block_y")) {
        // Validate the EFFECTIVE block shape (explicit values mixed with automatic ones) before
        // committing it; a rejected value leaves the context unchanged.
        const bool is_x = key[6] == 'x';
        const int lim = c->prm.dim == 2 ? 256 : LBM_STEP_MAX_THREADS;
        if (is_x && (value < 0 || value > lim || (value % 32)))
            return fail(LBM_ERR_INVALID, "block_x must be a multiple of 32 <= %d (0 = automatic)", lim);
        if (!is_x && (value < 0 || value > lim)) return fail(LBM_ERR_INVALID, "block_y must be in [0, %d] (0 = automatic)", lim);
        const int old_x = c->block_x, old_y = c->block_y;
        if (is_x) c->block_x = (int)value; else c->block_y = (int)value;
        if (c->prm.dim == 3) {
            dim3 blk, g;
            choose_block(c, &blk, &g);
            if (blk.x * blk.y > (unsigned)LBM_STEP_MAX_THREADS) {
                c->block_x = old_x; c->block_y = old_y;
                return fail(LBM_ERR_INVALID, "block %ux%u has more than %d threads", blk.x, blk.y, LBM_STEP_MAX_THREADS);
            }
        } else if (c->block_x > 0 && c->block_y > 0 && c->block_x * c->block_y > 256) {
            c->block_x = old_x; c->block_y = old_y;
            return fail(LBM_ERR_INVALID, "block_x*block_y > 256");
        }
    } else if (!strcmp(key, "pair")) {
        const bool was = pair_enabled(c);
        c->pair = value < 0 ? -1 : (value != 0);
        if (c->prm.nranks > 1 && pair_enabled(c) && !was) {
            // Fused pairs read two ghost planes per side and the neighbours' stale-f14 rows: fill them
            // now (on several GPUs switching `pair` on is therefore a COLLECTIVE call).
            CK(cudaSetDevice(c->dev));
            int rc = poll_flag(c);
            if (rc) return rc;
            if ((rc = halo_exchange(c, c->lat[c->cur], c->s_main, true, c->steps))) return rc;
            CK(cudaStreamSynchronize(c->s_main));
        }
        return LBM_OK;
    } else if (!strcmp(key, "pair_variant")) {
        if (value < -1 || value >= PAIR_NVARIANTS) return fail(LBM_ERR_INVALID, "pair_variant must be in [-1, %d)", PAIR_NVARIANTS);
        c->pair_variant = (int)value;
        return make_tensor_maps(c);
    } else if (!strcmp(key, "pair_lz")) {
        if (value < 0) return fail(LBM_ERR_INVALID, "pair_lz must be >= 0");
        c->pair_lz = (int)value;
        return LBM_OK;
    } else if (!strcmp(key, "halo_timeout_ms")) {
        if (value < 1) return fail(LBM_ERR_INVALID, "halo_timeout_ms must be >= 1");
        c->halo_timeout_ms = value;
        return LBM_OK;
    } else if (!strcmp(key, "converge_on_device")) {
        c->conv_mode = value != 0;
        return LBM_OK;
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
    // partial-sum buffer depends on the block shape
    if (c->prm.dim == 3 && c->prm.track_residual) return ensure_partials(c);
    return LBM_OK;
}

}  // extern "C"

