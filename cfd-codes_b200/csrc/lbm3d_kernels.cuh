// lbm3d_kernels.cuh -- sm_100a kernels of the D3Q19 cavity hot path (fp64, bit-exact vs the reference).
//
// Reference path: /root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c ("3D:")
//   macrocollideandstream 3D:215-321, HechtBC 3D:385-675, diff 3D:678-685, macrovars 3D:695-719,
//   calcvmag 3D:730-738, calcfeq/checkfeq 3D:163-169, 687-692.
//
// Formulation (DESIGN.md "Kernels"): ONE fused pull kernel per time step.  The lattice in HBM
// holds POST-collision populations S; a thread owns one cell x and
//   1. gathers F_q = S_old[x - c_q][q]            (what the reference's push streaming delivers),
//      w[q] where the source lies outside the box (slots the reference never writes),
//   2. applies the cell-local wall / lid / edge / corner rule of x's boundary class (d3q19_rules.h),
//      -> F is now bit-identical to the reference's f2[x][:] after HechtBC,
//   3. collides (BGK or TRT) in the reference's operation order and stores S_new[x][:] -- 19 aligned,
//      coalesced 8-byte stores; 19 loads, 19 stores, 304 algorithmic bytes per cell update.
// No tensor cores: the work is a stencil, not a contraction.  Compile with -fmad=false: FMA
// contraction changes the bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "d3q19_rules.h"

namespace lbm3d {

constexpr int Q = 19;
// 3D:31-33
#define LBM_W0 (1.0 / 3.0)
#define LBM_W1 (1.0 / 18.0)
#define LBM_W2 (1.0 / 36.0)

// X(q, cx, cy, cz, weight) -- velocities 3D:36-38, weights 3D:35
#define LBM_D3Q19_VEL(X)                                                                             \
    X(0, 0, 0, 0, LBM_W0)                                                                            \
    X(1, 1, 0, 0, LBM_W1) X(2, -1, 0, 0, LBM_W1) X(3, 0, 1, 0, LBM_W1) X(4, 0, -1, 0, LBM_W1)        \
    X(5, 0, 0, 1, LBM_W1) X(6, 0, 0, -1, LBM_W1)                                                     \
    X(7, 1, 1, 0, LBM_W2) X(8, -1, -1, 0, LBM_W2) X(9, 1, 0, 1, LBM_W2) X(10, -1, 0, -1, LBM_W2)     \
    X(11, 0, 1, 1, LBM_W2) X(12, 0, -1, -1, LBM_W2) X(13, 1, -1, 0, LBM_W2) X(14, -1, 1, 0, LBM_W2)  \
    X(15, 1, 0, -1, LBM_W2) X(16, -1, 0, 1, LBM_W2) X(17, 0, 1, -1, LBM_W2) X(18, 0, -1, 1, LBM_W2)

enum Mode : int {
    MODE_STEP = 0,              // gather + BC + collide + store S_new
    MODE_STEP_STOREF = 1,       // ... and also store F (the reference's f2) to the F buffer
    MODE_STEP_DIFF_STOREF = 2,  // ... and accumulate sum (F - Fbuf)^2 before overwriting Fbuf
    MODE_MATERIALISE = 3        // gather + BC, store F only (no collide, no state change)
};

// Kernel parameters (passed by value; lives in the constant bank).
struct KP {
    // Per-population base pointers, prepared by the host so that ONE 32-bit cell index addresses all
    // 38 accesses of a cell (one IMAD.WIDE each instead of 64-bit multiply chains):
    //   srcq[q][cell] = S_old[x - c_q][q]   (base already shifted by -(cx + cy*N + cz*N^2))
    //   dstq[q][cell] = S_new[x][q]
    // with cell = (kl+G)*N^2 + j*N + i (G ghost planes first).  The lattices are allocated with
    // N^2+N+1 elements of padding at both ends, so the shifted reads of box-boundary cells stay
    // inside the allocation (their values are discarded, see k_step).
    const double *srcq[19];
    double *dstq[19];
    double *Fbuf;        // F buffer [q][nzl*N2] (no ghosts), may be null for MODE_STEP
    const double *side_rd;  // stale f14 of the x=N-1 face, [nzl][N] (written two iterations ago)
    double *side_wr;        // slot this iteration writes
    const double *side_rd2; // the same for the second iteration of k_step_pair
    double *side_wr2;
    double *partials;    // per-block partial sums (MODE_STEP_DIFF_STOREF)
    int *flag;           // sticky negative-feq flag
    int N, M;            // grid points per side, N-1
    int Mz;              // global N-1 along z
    int nzl, k0;         // owned planes, global index of local plane 0
    int kl0, kstride;    // this launch covers local planes kl0 + blockIdx.z * kstride
    int G;               // ghost planes below local plane 0 (cell = (kl+G)*N^2 + j*N + i)
    int zlo, zhi, Lz;    // k_step_pair: output planes [zlo, zhi) of this launch, planes per z segment
    int nseg1, zlo2, zhi2;  // ... blocks with blockIdx.z >= nseg1 cover a second range [zlo2, zhi2) (slab-boundary launch)
    int partial_base;    // offset of this launch's block partials
    long long N2, PS, FS;
    double omega, omegam, omomega, gammainv;   // 3D:21,23,30,40
    double ulid, vlid, wlid, vfrac, third, sixth;  // 3D:183-191, 386, 41-42
    // Direct NVLink halo (PEER kernels only): the neighbours' ghost planes of the lattice being
    // written, mapped into this process (CUDA IPC / peer access); indexed by j*N + i.
    double *peer_dn[5];          // rank-1's top ghost plane for populations {6,10,12,15,17}; null at the bottom rank
    double *peer_up[5];          // rank+1's bottom ghost plane for {5,9,11,16,18}; null at the top rank
    unsigned int *done_count;    // blocks of this launch that have finished their peer stores
    long long *sig_dn, *sig_up;  // the neighbours' arrival flags (in THEIR memory), null if absent
    long long sig_value;         // iteration stamp to publish
};

// ---- global memory access policy (tuned with tools/ab_variants.sh; default = plain LDG/STG) ------
__device__ __forceinline__ double lbm_load(const double *p) {
#if defined(LBM_LD_NA)
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#elif defined(LBM_LD_CS)
    return __ldcs(p);
#elif defined(LBM_LD_CG)
    return __ldcg(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void lbm_store(double *p, double v) {
#if defined(LBM_ST_CS)
    __stcs(p, v);
#elif defined(LBM_ST_CG)
    __stcg(p, v);
#elif defined(LBM_ST_WT)
    __stwt(p, v);
#else
    *p = v;
#endif
}

// ---- collision --------------------------------------------------------------------------------
// In place: f (pre-collision, the reference's f1[x][:]) -> post-collision f*.  Returns true if any
// equilibrium was negative (checkfeq, 3D:687-692).
//
// Operation order follows 3D:224-246 (BGK) and 3D:268-302 (TRT) exactly; the only rewrites are
// exact identities (no rounding changes): multiplying by c in {-1,0,1} becomes add/sub/skip;
// cdotu(opp q) = -cdotu(q), so 3*cdotu and (4.5*cdotu^2 - 1.5*U2)*gammainv are shared by a pair;
// w_q*rho has three distinct values; axis pairs reuse u*u, v*v, w*w from U2; and scaling by 0.5
// commutes with rounding (no under/overflow at these magnitudes), so
//   OMEGA*(0.5*(a+b) - 0.5*(c+d)) == (0.5*OMEGA)*((a+b) - (c+d))   bit for bit (3D:293-299).
// checkfeq (feq < 0.0, 3D:688) is evaluated on the sign bits: the OR of the high words is
// negative iff some feq has its sign bit set (feq = -0.0 cannot occur for finite rho != 0).
// G1: GAMMA == 1.0 (the reference default): x * gammainv == x exactly, so the multiply is dropped.
// (collide and apply_boundary are __host__ __device__ only so that tests/host_device_functions_check.cu
//  can run the very same source on the CPU against the oracle; the library itself exports no host
//  compute path.)
__host__ __device__ __forceinline__ int lbm_hi32(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, sizeof b);
    return (int)(b >> 32);
#endif
}

template <int METHOD, bool G1>
__host__ __device__ __forceinline__ bool collide(double (&f)[Q], const KP &p) {
    // 3D:268-270 / 3D:225-235 (the += loop visits q = 0..18: same order, zero terms are exact)
    double rho = f[0] + f[1];
    rho += f[2]; rho += f[3]; rho += f[4]; rho += f[5]; rho += f[6]; rho += f[7]; rho += f[8];
    rho += f[9]; rho += f[10]; rho += f[11]; rho += f[12]; rho += f[13]; rho += f[14];
    rho += f[15]; rho += f[16]; rho += f[17]; rho += f[18];
    // 3D:272-281: x-moving 1,2,7,8,9,10,13,14,15,16; y: 3,4,7,8,11,12,13,14,17,18; z: 5,6,9,..,18
    double us = f[1] - f[2];
    us += f[7]; us -= f[8]; us += f[9]; us -= f[10]; us += f[13]; us -= f[14]; us += f[15]; us -= f[16];
    double vs = f[3] - f[4];
    vs += f[7]; vs -= f[8]; vs += f[11]; vs -= f[12]; vs -= f[13]; vs += f[14]; vs += f[17]; vs -= f[18];
    double ws = f[5] - f[6];
    ws += f[9]; ws -= f[10]; ws += f[11]; ws -= f[12]; ws -= f[15]; ws += f[16]; ws -= f[17]; ws += f[18];

    double u, v, w;
    if (METHOD == 1) {  // TRT multiplies by the reciprocal (3D:271-281)
        const double rinv = 1.0 / rho;
        u = us * rinv; v = vs * rinv; w = ws * rinv;
    } else {            // BGK divides (3D:237-239)
        u = us / rho; v = vs / rho; w = ws / rho;
    }
    const double uu = u * u, vv = v * v, ww = w * w;
    const double U2 = (uu + vv) + ww;   // 3D:240, 284
    const double k15 = 1.5 * U2;
    const double wr0 = LBM_W0 * rho, wr1 = LBM_W1 * rho, wr2 = LBM_W2 * rho;  // `w * rho * (...)`, 3D:166
    int negbits = 0;
    const double homega = 0.5 * p.omega, homegam = 0.5 * p.omegam;  // exact

    {   // rest population, cdotu = 0: feq = w*rho*(1.0 + (0 - 1.5*U2)*gammainv)
        const double fe = wr0 * (1.0 + (G1 ? (-k15) : (-k15) * p.gammainv));
        negbits |= lbm_hi32(fe);
        if (METHOD == 1) f[0] = f[0] - p.omega * (f[0] - fe);        // 3D:286
        else             f[0] = p.omomega * f[0] + p.omega * fe;     // 3D:246
    }
#define LBM_PAIR(q, c, cc, wr)                                                                       \
    {                                                                                                \
        const double t3 = 3.0 * (c);                                                                 \
        const double E = G1 ? (4.5 * (cc) - k15) : (4.5 * (cc) - k15) * p.gammainv;                  \
        const double fe1 = (wr) * ((1.0 + t3) + E);                                                  \
        const double fe2 = (wr) * ((1.0 - t3) + E);                                                  \
        negbits |= lbm_hi32(fe1) | lbm_hi32(fe2);                                             \
        if (METHOD == 1) { /* 3D:293-302 */                                                          \
            const double fplus = homega * ((f[q] + f[q + 1]) - (fe1 + fe2));                         \
            const double fminus = homegam * ((f[q] - f[q + 1]) - (fe1 - fe2));                       \
            const double a = f[q] - fplus - fminus;                                                  \
            const double b = f[q + 1] - fplus + fminus;                                              \
            f[q] = a;                                                                                \
            f[q + 1] = b;                                                                            \
        } else { /* 3D:246 */                                                                        \
            f[q] = p.omomega * f[q] + p.omega * fe1;                                                 \
            f[q + 1] = p.omomega * f[q + 1] + p.omega * fe2;                                         \
        }                                                                                            \
    }
    LBM_PAIR(1, u, uu, wr1)
    LBM_PAIR(3, v, vv, wr1)
    LBM_PAIR(5, w, ww, wr1)
    { const double c = u + v; LBM_PAIR(7, c, c * c, wr2) }
    { const double c = u + w; LBM_PAIR(9, c, c * c, wr2) }
    { const double c = v + w; LBM_PAIR(11, c, c * c, wr2) }
    { const double c = u - v; LBM_PAIR(13, c, c * c, wr2) }
    { const double c = u - w; LBM_PAIR(15, c, c * c, wr2) }
    { const double c = v - w; LBM_PAIR(17, c, c * c, wr2) }
#undef LBM_PAIR
    return negbits < 0;
}

// ---- wall / lid / edge / corner rules (HechtBC, 3D:385-675) --------------------------------------
#define LBM_OPP(q) (((q) & 1) ? (q) + 1 : (q) - 1)
#define LBM_EDGE_OP_Z(x) (x)
#define LBM_EDGE_OP_P(x) ((x) + temp)
#define LBM_EDGE_OP_M(x) ((x) - temp)

#define LBM_FACE_CASE(bx, by, bz, un, on, a0, a1, a2, b0, b1, b2, c0, c1, c2, d0, d1, d2, u0, o0, s0, u1, o1, \
                      s1, u2, o2, s2, u3, o3, s3)                                                             \
    case LBM_CLS(bx, by, bz): {                                                                               \
        f[un] = f[on];                                                                                        \
        const double n0 = 0.5 * ((f[a0] + f[a1] + f[a2]) - (f[b0] + f[b1] + f[b2]));                          \
        const double n1 = 0.5 * ((f[c0] + f[c1] + f[c2]) - (f[d0] + f[d1] + f[d2]));                          \
        f[u0] = f[o0] s0 n0;                                                                                  \
        f[u1] = f[o1] s1 n0;                                                                                  \
        f[u2] = f[o2] s2 n1;                                                                                  \
        f[u3] = f[o3] s3 n1;                                                                                  \
    } break;

#define LBM_EDGE_CASE(bx, by, bz, coef, ta, tb, u0, s0, u1, s1, u2, s2, u3, s3, u4, s4, u5, s5, u6, s6, u7, s7, \
                      u8, s8)                                                                                   \
    case LBM_CLS(bx, by, bz): {                                                                                 \
        const double temp = (coef) * (f[ta] - f[tb]);                                                           \
        f[u0] = LBM_EDGE_OP_##s0(f[LBM_OPP(u0)]);                                                               \
        f[u1] = LBM_EDGE_OP_##s1(f[LBM_OPP(u1)]);                                                               \
        f[u2] = LBM_EDGE_OP_##s2(f[LBM_OPP(u2)]);                                                               \
        f[u3] = LBM_EDGE_OP_##s3(f[LBM_OPP(u3)]);                                                               \
        f[u4] = LBM_EDGE_OP_##s4(f[LBM_OPP(u4)]);                                                               \
        f[u5] = LBM_EDGE_OP_##s5(f[LBM_OPP(u5)]);                                                               \
        f[u6] = LBM_EDGE_OP_##s6(f[LBM_OPP(u6)]);                                                               \
        f[u7] = LBM_EDGE_OP_##s7(f[LBM_OPP(u7)]);                                                               \
        f[u8] = LBM_EDGE_OP_##s8(f[LBM_OPP(u8)]);                                                               \
    } break;

#define LBM_CORNER_CASE(bx, by, bz, u0, u1, u2, u3, u4, u5) \
    case LBM_CLS(bx, by, bz): {                             \
        f[u0] = f[LBM_OPP(u0)];                             \
        f[u1] = f[LBM_OPP(u1)];                             \
        f[u2] = f[LBM_OPP(u2)];                             \
        f[u3] = f[LBM_OPP(u3)];                             \
        f[u4] = f[LBM_OPP(u4)];                             \
        f[u5] = f[LBM_OPP(u5)];                             \
    } break;

__host__ __device__ __forceinline__ void apply_boundary(double (&f)[Q], const int cls, const KP &p) {
    switch (cls) {
        LBM_D3Q19_FACES(LBM_FACE_CASE)
        case LBM_CLS(0, 2, 0): {  // the moving lid y = N-1, 3D:437-445
            const double rho =
                p.vfrac * (f[1] + f[2] + f[5] + f[6] + f[9] + f[15] + f[16] + f[10] + f[0] +
                           2.0 * (f[3] + f[7] + f[14] + f[11] + f[17]));
            f[4] = f[3] - rho * p.vlid * p.third;
            const double nyx = 0.5 * ((f[1] + f[9] + f[15]) - (f[2] + f[16] + f[10])) - rho * p.ulid * p.third;
            const double nyz = 0.5 * ((f[5] + f[9] + f[16]) - (f[6] + f[15] + f[10])) - rho * p.wlid * p.third;
            f[8] = f[7] - rho * (p.vlid + p.ulid) * p.sixth + nyx;
            f[13] = f[14] + rho * (-p.vlid + p.ulid) * p.sixth - nyx;
            f[12] = f[11] - rho * (p.vlid + p.wlid) * p.sixth + nyz;
            f[18] = f[17] + rho * (-p.vlid + p.wlid) * p.sixth - nyz;
        } break;
        LBM_D3Q19_EDGES(LBM_EDGE_CASE)
        LBM_D3Q19_CORNERS(LBM_CORNER_CASE)
        default: break;
    }
}

// ---- block reduction (deterministic: fixed shuffle tree, fixed warp order) -------------------------
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[32];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) warp_part[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (tid == 0) {
        const int nw = (nthreads + 31) >> 5;
        for (int w = 0; w < nw; w++) s += warp_part[w];
    }
    return s;  // valid in thread 0
}

// ---- the fused step kernel ------------------------------------------------------------------------
// grid = (ceil(N/bx), ceil(N/by), planes of this launch); block = (bx, by), bx*by <= 128.
// Launch bounds tuned on B200 (profiles/r01_ab_occupancy.log): 128-thread blocks, >= 5 resident
// blocks per SM (<= 102 registers; ptxas uses 94 without spilling) -> 20 warps/SM.  With the looser
// bound (122 registers, 16 warps/SM) the kernel matches the copy bandwidth only at boost clocks and
// loses 2.5 % once the 1000 W power cap pulls the SM clock to ~1.6 GHz; a tighter bound spills.
#ifndef LBM_STEP_MAX_THREADS
#define LBM_STEP_MAX_THREADS 128
#define LBM_STEP_MIN_BLOCKS 5
#endif
// PEER (slab-boundary planes of a multi-GPU run with the direct NVLink halo): the thread also stores
// the five face-crossing post-collision populations of its cell straight into the neighbour GPU's
// ghost plane (peer memory), and the last block to finish publishes the iteration stamp in the
// neighbours' arrival flags -- compute and exchange in ONE kernel, no send/recv.
template <int METHOD, int MODE, bool G1, bool PEER = false>
__global__ void __launch_bounds__(LBM_STEP_MAX_THREADS, LBM_STEP_MIN_BLOCKS) k_step(const __grid_constant__ KP p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = p.kl0 + blockIdx.z * p.kstride;
    const bool active = (i < p.N) && (j < p.N);
    double dsum = 0.0;

    if (active) {
        const int kg = kl + p.k0;
        const int bx = (i == 0) ? 1 : ((i == p.M) ? 2 : 0);
        const int by = (j == 0) ? 1 : ((j == p.M) ? 2 : 0);
        const int bz = (kg == 0) ? 1 : ((kg == p.Mz) ? 2 : 0);
        const int cls = LBM_CLS(bx, by, bz);
        const int N1 = p.N;
        const int rowcell = j * N1 + i;
        const int cell = (kl + p.G) * (int)p.N2 + rowcell;   // < 2^31 (checked by the host)

        double f[Q];
        // 1. pull: unconditional, coalesced 8-byte loads through the pre-shifted bases.
#define LBM_PULL(q, cx, cy, cz, wq) f[q] = lbm_load(p.srcq[q] + cell);
        LBM_D3Q19_VEL(LBM_PULL)
#undef LBM_PULL

        // 2. boundary cells only
        if (cls != 0) {
            // A source outside the box: the reference never writes that slot, which therefore still
            // holds w[q] (buried slots) or is overwritten by the rule below.  The value
            // loaded above (a wrapped neighbour or padding) is discarded.
#define LBM_FIX(q, cx, cy, cz, wq)                                                                   \
            if ((cx == 1 && bx == 1) || (cx == -1 && bx == 2) || (cy == 1 && by == 1) ||             \
                (cy == -1 && by == 2) || (cz == 1 && bz == 1) || (cz == -1 && bz == 2))              \
                f[q] = (wq);
            LBM_D3Q19_VEL(LBM_FIX)
#undef LBM_FIX
            const bool xm_face = (cls == LBM_CLS(2, 0, 0));
            // stale f14 (3D:421): at x = N-1 the rule reads population 14 before assigning it; the
            // reference finds there what HechtBC wrote into the same buffer two iterations ago.
            if (xm_face) f[14] = p.side_rd[kl * N1 + j];
            apply_boundary(f, cls, p);
            if (MODE != MODE_MATERIALISE) {
                if (xm_face) p.side_wr[kl * N1 + j] = f[14];
            }
        }

        // f == the reference's f2[x][:] after HechtBC of this iteration
        if (MODE != MODE_STEP) {
            double *F = p.Fbuf + ((long long)kl * p.N2 + rowcell);
            if (MODE == MODE_STEP_DIFF_STOREF) {
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const double d = f[q] - F[(long long)q * p.FS];
                    dsum += d * d;   // diff(), 3D:682 (the reference's summation order is unspecified)
                }
            }
#pragma unroll
            for (int q = 0; q < Q; q++) F[(long long)q * p.FS] = f[q];
        }

        // 3. collide + store
        if (MODE != MODE_MATERIALISE) {
            const bool neg = collide<METHOD, G1>(f, p);
            if (neg) atomicOr(p.flag, 1);
#pragma unroll
            for (int q = 0; q < Q; q++) lbm_store(p.dstq[q] + cell, f[q]);
            if (PEER) {
                if (kl == 0 && p.peer_dn[0] != nullptr) {          // cz = -1 populations leave through the bottom face
                    p.peer_dn[0][rowcell] = f[6];  p.peer_dn[1][rowcell] = f[10]; p.peer_dn[2][rowcell] = f[12];
                    p.peer_dn[3][rowcell] = f[15]; p.peer_dn[4][rowcell] = f[17];
                }
                if (kl == p.nzl - 1 && p.peer_up[0] != nullptr) {  // cz = +1 populations leave through the top face
                    p.peer_up[0][rowcell] = f[5];  p.peer_up[1][rowcell] = f[9];  p.peer_up[2][rowcell] = f[11];
                    p.peer_up[3][rowcell] = f[16]; p.peer_up[4][rowcell] = f[18];
                }
            }
        }
    }

    if (PEER) {
        // publish: every block fences its peer stores system-wide, the last one raises the flags
        __shared__ bool last;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
            last = (atomicAdd(p.done_count, 1u) == total - 1);
        }
        __syncthreads();
        if (last && threadIdx.x == 0 && threadIdx.y == 0) {
            *p.done_count = 0;
            __threadfence_system();
            if (p.sig_dn) *reinterpret_cast<volatile long long *>(p.sig_dn) = p.sig_value;
            if (p.sig_up) *reinterpret_cast<volatile long long *>(p.sig_up) = p.sig_value;
            __threadfence_system();
        }
    }

    if (MODE == MODE_STEP_DIFF_STOREF) {
        const double s = block_sum(dsum);
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            p.partials[p.partial_base + bid] = s;
        }
    }
}

// ---- 128-bit variant of the plain step: two x-adjacent cells per thread --------------------------
// (north_star item 2.)  Needs N even.  The 9 populations with cx = 0 are pulled with one aligned
// 16-byte load per pair of cells, the 10 x-shifted ones with two 8-byte loads (their pair straddles
// a 16-byte boundary); all 19 are stored with aligned 16-byte stores.  Same arithmetic, same bits.
#ifndef LBM_STEP2_MIN_BLOCKS
#define LBM_STEP2_MIN_BLOCKS 3
#endif
template <int METHOD, bool G1>
__global__ void __launch_bounds__(128, LBM_STEP2_MIN_BLOCKS) k_step_x2(const __grid_constant__ KP p) {
    const int i0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = p.kl0 + blockIdx.z * p.kstride;
    if (i0 >= p.N || j >= p.N) return;
    const int kg = kl + p.k0;
    const int by = (j == 0) ? 1 : ((j == p.M) ? 2 : 0);
    const int bz = (kg == 0) ? 1 : ((kg == p.Mz) ? 2 : 0);
    const int bxa = (i0 == 0) ? 1 : 0;                 // cell a = (i0, j, k): can only touch x = 0
    const int bxb = (i0 + 1 == p.M) ? 2 : 0;           // cell b = (i0+1, j, k): can only touch x = N-1
    const int clsa = LBM_CLS(bxa, by, bz), clsb = LBM_CLS(bxb, by, bz);
    const int N1 = p.N;
    const int cell = (kl + p.G) * (int)p.N2 + j * N1 + i0;

    double fa[Q], fb[Q];
#define LBM_PULL2(q, cx, cy, cz, wq)                                                                 \
    if (cx == 0) {                                                                                   \
        const double2 v = *reinterpret_cast<const double2 *>(p.srcq[q] + cell);                      \
        fa[q] = v.x; fb[q] = v.y;                                                                    \
    } else {                                                                                         \
        fa[q] = p.srcq[q][cell]; fb[q] = p.srcq[q][cell + 1];                                        \
    }
    LBM_D3Q19_VEL(LBM_PULL2)
#undef LBM_PULL2

#define LBM_FIXA(q, cx, cy, cz, wq)                                                                  \
    if ((cx == 1 && bxa == 1) || (cy == 1 && by == 1) || (cy == -1 && by == 2) || (cz == 1 && bz == 1) || \
        (cz == -1 && bz == 2))                                                                       \
        fa[q] = (wq);
#define LBM_FIXB(q, cx, cy, cz, wq)                                                                  \
    if ((cx == -1 && bxb == 2) || (cy == 1 && by == 1) || (cy == -1 && by == 2) || (cz == 1 && bz == 1) || \
        (cz == -1 && bz == 2))                                                                       \
        fb[q] = (wq);
    if (clsa != 0) {
        LBM_D3Q19_VEL(LBM_FIXA)
        apply_boundary(fa, clsa, p);
    }
    if (clsb != 0) {
        LBM_D3Q19_VEL(LBM_FIXB)
        const bool xm_face = (clsb == LBM_CLS(2, 0, 0));
        if (xm_face) fb[14] = p.side_rd[kl * N1 + j];   // stale f14, see k_step
        apply_boundary(fb, clsb, p);
        if (xm_face) p.side_wr[kl * N1 + j] = fb[14];
    }
#undef LBM_FIXA
#undef LBM_FIXB
    const bool nega = collide<METHOD, G1>(fa, p);
    const bool negb = collide<METHOD, G1>(fb, p);
    if (nega | negb) atomicOr(p.flag, 1);
#pragma unroll
    for (int q = 0; q < Q; q++) *reinterpret_cast<double2 *>(p.dstq[q] + cell) = make_double2(fa[q], fb[q]);
}

// S = collide(F) for every owned cell, F taken from the F buffer (restart, lbm_set_f), or
// F = w (initial state, 3D:192-205: feq(rho=1,u=0) = w[q] exactly) when Fbuf == nullptr.
// Covers ghost planes too when `with_ghosts` (initial state only: uniform).
template <int METHOD>
__global__ void __launch_bounds__(128, 4) k_collide_cells(const __grid_constant__ KP p, const int with_ghosts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int kl = (int)blockIdx.z - (with_ghosts ? p.G : 0);
    if (i >= p.N || j >= p.N) return;
    const long long rowcell = (long long)j * p.N + i;
    double f[Q];
    if (p.Fbuf != nullptr && !with_ghosts) {
        const double *F = p.Fbuf + ((long long)kl * p.N2 + rowcell);
#pragma unroll
        for (int q = 0; q < Q; q++) f[q] = F[(long long)q * p.FS];
    } else {
#define LBM_SETW(q, cx, cy, cz, wq) f[q] = (wq);
        LBM_D3Q19_VEL(LBM_SETW)
#undef LBM_SETW
    }
    const bool neg = collide<METHOD, false>(f, p);
    if (neg) atomicOr(p.flag, 1);
    const int cell = (kl + p.G) * (int)p.N2 + (int)rowcell;
#pragma unroll
    for (int q = 0; q < Q; q++) p.dstq[q][cell] = f[q];
}

// Fill the F buffer with the initial distributions f = w[q] (3D:192-205).
__global__ void k_fill_w(double *F, long long FS) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= FS) return;
#define LBM_FILLW(q, cx, cy, cz, wq) F[(long long)(q) * FS + n] = (wq);
    LBM_D3Q19_VEL(LBM_FILLW)
#undef LBM_FILLW
}

// out[r] = plane[r*N + N-1]: the x = N-1 column of one population over all (k, j) rows (lbm_set_f:
// rebuilds the stale-f14 history from an uploaded state).
__global__ void k_gather_xmax(const double *__restrict__ plane, int N, long long rows, double *__restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) out[r] = plane[r * N + (N - 1)];
}

__global__ void k_fill(double *a, long long n, double v) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] = v;
}

// macrovars (3D:695-719) + calcvmag (3D:730-738) from the F buffer.
__global__ void __launch_bounds__(256) k_macrovars(const double *__restrict__ F, long long FS, double *u1,
                                                   double *u2, double *u3, double *rho, double *vmag) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= FS) return;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = F[(long long)q * FS + n];
    // `+=` loops over q = 0..18 starting from 0.0 (3D:701-711); zero terms are exact
    double r = 0.0 + f[0];
#pragma unroll
    for (int q = 1; q < Q; q++) r += f[q];
    double us = 0.0 + f[1];
    us -= f[2]; us += f[7]; us -= f[8]; us += f[9]; us -= f[10]; us += f[13]; us -= f[14]; us += f[15]; us -= f[16];
    double vs = 0.0 + f[3];
    vs -= f[4]; vs += f[7]; vs -= f[8]; vs += f[11]; vs -= f[12]; vs -= f[13]; vs += f[14]; vs += f[17]; vs -= f[18];
    double ws = 0.0 + f[5];
    ws -= f[6]; ws += f[9]; ws -= f[10]; ws += f[11]; ws -= f[12]; ws -= f[15]; ws += f[16]; ws -= f[17]; ws += f[18];
    const double u = us / r, v = vs / r, w = ws / r;   // 3D:713-715 divides
    if (rho) rho[n] = r;
    if (u1) u1[n] = u;
    if (u2) u2[n] = v;
    if (u3) u3[n] = w;
    if (vmag) vmag[n] = sqrt((u * u + v * v) + w * w);  // 3D:735
}

// Wait (one thread) until both neighbours have published `need` in this GPU's arrival flags, i.e.
// their slab-boundary kernels of the previous iteration -- the ones that fill my ghost planes and
// read the ghost planes I am about to overwrite -- are complete.  Bounded in WALL-CLOCK time
// (%globaltimer, nanoseconds; independent of the SM clock): after `timeout_ns` the error bit 4 is
// raised instead of hanging the GPU.  Ranks must therefore stay within that skew of each other
// (include/lbm_b200.h, option halo_timeout_ms).
__device__ __forceinline__ unsigned long long lbm_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_wait_arrivals(const volatile long long *sig, long long need_dn, long long need_up, int *flag,
                                unsigned long long timeout_ns) {
    const unsigned long long t0 = lbm_globaltimer_ns();
    while (sig[0] < need_dn || sig[1] < need_up) {
        if (lbm_globaltimer_ns() - t0 > timeout_ns) { atomicOr(flag, 4); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// Second stage of the residual: one block sums the block partials in a fixed order.
__global__ void __launch_bounds__(1024) k_sum_partials(const double *__restrict__ part, long long n, double *out) {
    __shared__ double sm[1024];
    double s = 0.0;
    for (long long t = threadIdx.x; t < n; t += 1024) s += part[t];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

}  // namespace lbm3d
