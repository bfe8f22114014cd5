// lbm3d_pair.cuh -- TWO time steps per pass over HBM (temporal blocking through shared memory).
//
// Reference loop being blocked: 3D:104-124 -- per iteration macrocollideandstream() (3D:215-321) +
// HechtBC(f2) (3D:385-675) + swap; two iterations of it read and write the whole lattice twice
// (2 x 304 B per cell).  This kernel produces S(t+2) from S(t) with ONE read and ONE write of the
// lattice (304 B per cell per TWO updates = 152 B per update) and the very same per-cell arithmetic
// as k_step (same collide<> / apply_boundary<>, same order), so the bits do not change.
//
// Decomposition.  A thread block owns a column of the box: an x-y tile of 30 x (TY-2) OUTPUT cells
// and a segment [z0, z1) of z planes, and marches along z.  Its 32 x TY threads cover the tile grown
// by one cell on every side (one warp per x-row).  Per z iteration `pz`:
//   step A (iteration t of the reference, all 32 x TY threads inside the box): exactly k_step's body for
//          cell (i, j, pz) -- pull 19 populations of S(t) from HBM, wall/lid rule, collide -- but the
//          result S(t+1) goes to SHARED MEMORY (16 populations) / registers (the 3 populations with
//          cx = cy = 0, which the same thread consumes) instead of HBM;
//   step B (iteration t+1, the 30 x (TY-2) inner threads): the same update for cell (i, j, pz-1),
//          pulling S(t+1) from shared memory: the up-going populations of plane pz-2, the in-plane
//          ones of plane pz-1 and the down-going ones of plane pz; the result S(t+2) goes to HBM.
// The halo ring of S(t+1) (1 cell in x and y, 1 plane at both ends of the z segment) is computed
// redundantly by the neighbouring blocks -- no communication between blocks, no global intermediate.
//
// Shared memory per thread-cell: up-going {9,11,16,18} x 3 planes, in-plane {1,2,3,4,7,8,13,14} x 2
// planes, down-going {10,12,15,17} x 1 plane = 32 doubles = 256 B  (TY = 10: 81 920 B per block, two
// blocks per SM; the staged variants add 152 B per cell).  Populations 0, 5, 6 never leave the thread.
//
// Stale f14 (3D:421): iteration t reads slot (t+2)%4 of the `side` history and writes slot t%4, so
// the two fused iterations read {t+2, t+3} and write {t, t+1} (mod 4): redundant step-A work in
// neighbouring blocks can neither race with step B nor see its output.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "lbm3d_kernels.cuh"

namespace lbm3d {

constexpr int PAIR_TX = 32;       // threads per tile row (one warp)
constexpr int PAIR_OX = 30;       // output cells per tile row

constexpr int PAIR_BOX_X = 34;    // TMA box width: the tile's 32 columns + one on either side
// doubles per population in the TMA staging area (box rows are dense; every box 128-byte aligned)
__host__ __device__ constexpr int pair_tma_block(int TY) { return ((TY * PAIR_BOX_X * 8 + 127) / 128) * 16; }
__host__ __device__ constexpr int pair_smem_bytes(int TY, int STAGE) {
    return 32 * PAIR_TX * TY * (int)sizeof(double) +
           (STAGE == 1 ? 19 * PAIR_TX * TY * (int)sizeof(double) : 0) + (STAGE == 4 ? 19 * pair_tma_block(TY) * (int)sizeof(double) + 16 : 0);
}

#define LBM_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))

// Variants (template parameters, A/B'd on the GPU: profiles/r02_summary.md):
//   MINB  resident blocks per SM the register budget is sized for (1: big tiles; 2: two smaller blocks
//         whose load and compute phases interleave);
//   STAGE 0: step A pulls straight from HBM/L2 into registers (L2 prefetch one plane ahead);
//         1: the 19 pulls of plane pz+1 are issued as asynchronous copies (cp.async, LDGSTS) into a
//            per-thread staging slot of shared memory right after the first barrier and land while
//            step B computes, so step A never waits for HBM (costs 152 B of shared memory per cell);
//         4: the same staging done by the TMA unit: ONE thread issues 19 tensor copies
//            (cp.async.bulk.tensor.4d, one 34 x TY box per population, rows and plane shifted by
//            -(cy, cz), out-of-range elements zero-filled) that complete on an mbarrier -- no
//            per-thread address arithmetic, no load instructions, nothing in the LSU queue.  A box must
//            start on a 16-byte boundary (measured: tools/tma_probe.cu), so every box starts at the
//            even column 30*bx - 2 and is two columns wider than the tile; the x shift -cx is applied
//            when the thread reads its value back.  Needs N even (16-byte row stride);
//         2: like 0 but software-pipelined through 19 extra registers (slower: 128 registers leave
//            16 warps per SM and the pulls still stall -- kept for the record);
//         3: 0 without the L2 prefetch (the 19 prefetch instructions per thread cost more than they save).
//
// grid = (ceil(N/30), ceil(N/(TY-2)), z segments); block = (32, TY); dynamic smem = pair_smem_bytes(TY, STAGE).
template <int METHOD, bool G1, int TY, int MINB, int STAGE>
__global__ void __launch_bounds__(PAIR_TX *TY, MINB) k_step_pair(const __grid_constant__ KP p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double sm[];
    constexpr int NT = PAIR_TX * TY;
    double *const s_up = sm;             // [3 planes][4][NT]   populations 9, 11, 16, 18
    double *const s_in = sm + 12 * NT;   // [2 planes][8][NT]   populations 1, 2, 3, 4, 7, 8, 13, 14
    double *const s_dn = sm + 28 * NT;   // [4][NT]             populations 10, 12, 15, 17
    double *const s_st = sm + 32 * NT;   // STAGE 1: [19][NT] this thread's pulls of the next plane; STAGE 4: [19][TY][34] TMA boxes
    constexpr int SB = pair_tma_block(TY);
    constexpr bool PIPE = (STAGE == 2);

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int li = ty * PAIR_TX + tx;
    const int i = (int)blockIdx.x * PAIR_OX - 1 + tx;
    const int j = (int)blockIdx.y * (TY - 2) - 1 + ty;
    const int N1 = p.N;
    const bool in_box = (i >= 0) && (i < N1) && (j >= 0) && (j < N1);
    const bool out_cell = in_box && (tx >= 1) && (tx <= PAIR_OX) && (ty >= 1) && (ty <= TY - 2);
    int zb = (int)blockIdx.z, zlo = p.zlo, zhi = p.zhi;
    if (zb >= p.nseg1) { zb -= p.nseg1; zlo = p.zlo2; zhi = p.zhi2; }   // second z range of a slab-boundary launch
    const int z0 = zlo + zb * p.Lz;
    const int z1 = min(z0 + p.Lz, zhi);
    const int bx = (i == 0) ? 1 : ((i == p.M) ? 2 : 0);
    const int by = (j == 0) ? 1 : ((j == p.M) ? 2 : 0);
    const int rowcell = j * N1 + i;
    const int N2i = (int)p.N2;

    double c0 = 0.0, c5a = 0.0, c5b = 0.0;    // S(t+1) populations 0 (previous plane) and 5 (two planes back)
    int su = 0;                                // s_up slot of plane pz (rotates 0,1,2)
    int cell = (z0 - 1 + p.G) * N2i + rowcell; // index of (i, j, pz) incl. ghost planes; < 2^31 (host-checked)

#define LBM_PULL_INTO(arr, at) { _Pragma("unroll") for (int q_ = 0; q_ < Q; q_++) arr[q_] = lbm_load(p.srcq[q_] + (at)); }
#define LBM_PLANE_IN_BOX(z) ((z) + p.k0 >= 0 && (z) + p.k0 <= p.Mz)
#define LBM_STAGE_PLANE(at)                                                                                         \
    {                                                                                                               \
        const unsigned dst_ = (unsigned)__cvta_generic_to_shared(s_st + li);                                        \
        _Pragma("unroll") for (int q_ = 0; q_ < Q; q_++)                                                            \
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_ + q_ * NT * 8), "l"(p.srcq[q_] + (at)) : "memory"); \
    }
    // STAGE 4: TMA.  Box (34, TY, 1, 1) of population q at lattice coordinates (30*bx - 2, y - cy, z - cz, q).
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(sm + 32 * NT + 19 * SB);
    const unsigned st_base = (unsigned)__cvta_generic_to_shared(s_st);
    const bool tma_leader = (tx == 0 && ty == 0);
    const int tile_x = (int)blockIdx.x * PAIR_OX - 1, tile_y = (int)blockIdx.y * (TY - 2) - 1;
    unsigned tma_phase = 0;
#define LBM_TMA_ONE(q, cx, cy, cz, wq)                                                                                     \
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" \
                 ::"r"(st_base + (q) * SB * 8), "l"(&tmap), "r"(mbar), "r"(tile_x - 1), "r"(tile_y - (cy)), "r"(zc_ - (cz)), "r"(q) \
                 : "memory");
#define LBM_TMA_PLANE(zplane)                                                                                              \
    {                                                                                                                      \
        const int zc_ = (zplane) + p.G;                                                                                    \
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(19 * TY * PAIR_BOX_X * 8) : "memory"); \
        LBM_D3Q19_VEL(LBM_TMA_ONE)                                                                                         \
    }
    if (STAGE == 4) {
        if (tma_leader) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the init must be visible to the TMA unit
        }
        __syncthreads();
        if (tma_leader && LBM_PLANE_IN_BOX(z0 - 1)) LBM_TMA_PLANE(z0 - 1)
    }
    double fn[Q];                              // PIPE: the pulls of the next plane, in flight
    if (PIPE) {
        if (in_box && LBM_PLANE_IN_BOX(z0 - 1)) LBM_PULL_INTO(fn, cell)
    }
    if (STAGE == 1) {
        if (in_box && LBM_PLANE_IN_BOX(z0 - 1)) LBM_STAGE_PLANE(cell)
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    for (int pz = z0 - 1; pz <= z1; pz++, cell += N2i) {
        const int kg = pz + p.k0;
        const bool have_a = in_box && (kg >= 0) && (kg <= p.Mz);
        double f[Q];
        f[0] = 0.0; f[5] = 0.0; f[6] = 0.0;
        // ---- step A: S(t) from HBM -> S(t+1) into shared memory --------------------------------------
        if (STAGE == 4 && (kg >= 0) && (kg <= p.Mz)) {   // the staged plane has landed (phase parity of the mbarrier)
            asm volatile(
                "{\n\t.reg .pred P1;\n\tLBM_TMA_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                "@P1 bra LBM_TMA_DONE;\n\tbra LBM_TMA_WAIT;\n\tLBM_TMA_DONE:\n\t}" ::"r"(mbar), "r"(tma_phase) : "memory");
            tma_phase ^= 1;
        }
        if (have_a) {
            if (PIPE) {
#pragma unroll
                for (int q = 0; q < Q; q++) f[q] = fn[q];
            } else if (STAGE == 4) {
                const double *b = s_st + ty * PAIR_BOX_X + tx + 1;     // column tx of the tile = column tx+1 of the box
#define LBM_UNBOX(q, cx, cy, cz, wq) f[q] = b[(q) * SB - (cx)];
                LBM_D3Q19_VEL(LBM_UNBOX)
#undef LBM_UNBOX
            } else if (STAGE == 1) {
                asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
                for (int q = 0; q < Q; q++) f[q] = s_st[q * NT + li];   // this thread's own copies, no barrier needed
            } else {
                LBM_PULL_INTO(f, cell)
            }
            const int bz = (kg == 0) ? 1 : ((kg == p.Mz) ? 2 : 0);
            const int cls = LBM_CLS(bx, by, bz);
            if (cls != 0) {
#define LBM_FIX(q, cx, cy, cz, wq)                                                                   \
                if ((cx == 1 && bx == 1) || (cx == -1 && bx == 2) || (cy == 1 && by == 1) ||         \
                    (cy == -1 && by == 2) || (cz == 1 && bz == 1) || (cz == -1 && bz == 2))          \
                    f[q] = (wq);
                LBM_D3Q19_VEL(LBM_FIX)
#undef LBM_FIX
                const bool xm_face = (cls == LBM_CLS(2, 0, 0));
                if (xm_face) f[14] = p.side_rd[pz * N1 + j];
                apply_boundary(f, cls, p);
                if (xm_face) p.side_wr[pz * N1 + j] = f[14];
            }
            if (collide<METHOD, G1>(f, p)) atomicOr(p.flag, 1);
            double *u = s_up + su * (4 * NT) + li;
            u[0 * NT] = f[9]; u[1 * NT] = f[11]; u[2 * NT] = f[16]; u[3 * NT] = f[18];
            double *n = s_in + (pz & 1) * (8 * NT) + li;
            n[0 * NT] = f[1]; n[1 * NT] = f[2]; n[2 * NT] = f[3]; n[3 * NT] = f[4];
            n[4 * NT] = f[7]; n[5 * NT] = f[8]; n[6 * NT] = f[13]; n[7 * NT] = f[14];
            double *d = s_dn + li;
            d[0 * NT] = f[10]; d[1 * NT] = f[12]; d[2 * NT] = f[15]; d[3 * NT] = f[17];
        }
        __syncthreads();
        if (PIPE) {
            // pulls of the next plane: in flight while step B computes
            if (in_box && pz < z1 && LBM_PLANE_IN_BOX(pz + 1)) LBM_PULL_INTO(fn, cell + N2i)
        }
        if (STAGE == 1) {
            if (in_box && pz < z1 && LBM_PLANE_IN_BOX(pz + 1)) LBM_STAGE_PLANE(cell + N2i)
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (STAGE == 4) {   // every thread has consumed the staged plane (barrier above): refill it
            if (tma_leader && pz < z1 && LBM_PLANE_IN_BOX(pz + 1)) LBM_TMA_PLANE(pz + 1)
        }
        // Warm L2 with the lines a later step A pulls (every (population, plane) pair is read exactly
        // once per block, so without this every pull is a DRAM-latency miss).
        if (STAGE != 1 && STAGE != 3 && STAGE != 4 && in_box && pz + (PIPE ? 1 : 0) < z1) {   // STAGE 3: variant 0 without the prefetch (A/B only)
            const int ahead = cell + (PIPE ? 2 : 1) * N2i;
#define LBM_PF(q, cx, cy, cz, wq) LBM_PREFETCH_L2(p.srcq[q] + ahead);
            LBM_D3Q19_VEL(LBM_PF)
#undef LBM_PF
        }
        // ---- step B: S(t+1) from shared memory -> S(t+2) into HBM, one plane behind -------------------
        const int P = pz - 1;
        if (out_cell && P >= z0) {
            double g[Q];
            const int su2 = (su == 2) ? 0 : su + 1;   // slot of plane pz-2 == (pz+1) mod 3
            const double *u = s_up + su2 * (4 * NT) + li;
            const double *n = s_in + (P & 1) * (8 * NT) + li;
            const double *d = s_dn + li;
            g[0] = c0; g[5] = c5a; g[6] = f[6];
            g[1] = n[0 * NT - 1];            g[2] = n[1 * NT + 1];
            g[3] = n[2 * NT - PAIR_TX];      g[4] = n[3 * NT + PAIR_TX];
            g[7] = n[4 * NT - 1 - PAIR_TX];  g[8] = n[5 * NT + 1 + PAIR_TX];
            g[13] = n[6 * NT - 1 + PAIR_TX]; g[14] = n[7 * NT + 1 - PAIR_TX];
            g[9] = u[0 * NT - 1];            g[11] = u[1 * NT - PAIR_TX];
            g[16] = u[2 * NT + 1];           g[18] = u[3 * NT + PAIR_TX];
            g[10] = d[0 * NT + 1];           g[12] = d[1 * NT + PAIR_TX];
            g[15] = d[2 * NT - 1];           g[17] = d[3 * NT - PAIR_TX];
            const int kg2 = kg - 1;
            const int bz = (kg2 == 0) ? 1 : ((kg2 == p.Mz) ? 2 : 0);
            const int cls = LBM_CLS(bx, by, bz);
            if (cls != 0) {
#define LBM_FIXB(q, cx, cy, cz, wq)                                                                  \
                if ((cx == 1 && bx == 1) || (cx == -1 && bx == 2) || (cy == 1 && by == 1) ||         \
                    (cy == -1 && by == 2) || (cz == 1 && bz == 1) || (cz == -1 && bz == 2))          \
                    g[q] = (wq);
                LBM_D3Q19_VEL(LBM_FIXB)
#undef LBM_FIXB
                const bool xm_face = (cls == LBM_CLS(2, 0, 0));
                if (xm_face) g[14] = p.side_rd2[P * N1 + j];
                apply_boundary(g, cls, p);
                if (xm_face) p.side_wr2[P * N1 + j] = g[14];
            }
            if (collide<METHOD, G1>(g, p)) atomicOr(p.flag, 1);
            const int cell2 = cell - N2i;
#pragma unroll
            for (int q = 0; q < Q; q++) lbm_store(p.dstq[q] + cell2, g[q]);
        }
        c5a = c5b; c5b = f[5]; c0 = f[0];
        su = (su == 2) ? 0 : su + 1;
        __syncthreads();
    }
#undef LBM_PULL_INTO
#undef LBM_STAGE_PLANE
#undef LBM_TMA_PLANE
#undef LBM_TMA_ONE
#undef LBM_PLANE_IN_BOX
}

}  // namespace lbm3d
