// lbm2d_kernels.cuh -- D2Q9 variant of the fused pull kernel (fp64), BASELINE config 5.
//
// Physics, constants and arithmetic follow the reference's 2D program
//   /root/reference/2D Lid-driven Cavity with Lattice Boltzmann Method/main.cpp   ("2D:")
// streamCollide 2D:281-416 (pull + half-way bounce-back walls 2D:308-344 + moving lid 2D:352-356,
// incompressible "deviation" form: the lattice stores f - w, no division by density 2D:360-363,
// TRT 2D:367-413), macroVars 2D:419-449.
//
// One deliberate difference: the reference updates ONE lattice in place, in sweep
// order, so a cell sees already-updated values from (i-1,.) and (i,j-1) and its result depends on
// the sweep and on the number of MPI ranks.  This kernel is the time-consistent
// two-lattice version of the same update: every cell reads the previous step.  Both have the same
// fixed point.  Layout: population-major [9][rows+2][N] (one ghost row below and above the slab).
#pragma once
#include <cuda_runtime.h>

namespace lbm2d {

constexpr int Q = 9;
// 2D:27-29
#define LBM2_W0 (4.0 / 9.0)
#define LBM2_WS (1.0 / 9.0)
#define LBM2_WD (1.0 / 36.0)

struct KP2 {
    const double *srcq[9];  // pre-shifted: srcq[k][cell] = dist_old[k](i - cx, j - cy)   (pulls 2D:292-348)
    const double *own[9];   // un-shifted:  own[k][cell]  = dist_old[k](i, j)             (wall bounce 2D:308-356)
    double *dstq[9];
    double *u, *v, *rho;    // macroVars state [rows][N] (2D:132-136)
    double *vmag;           // output path only: sqrt(u^2 + v^2), may be null
    double *partials;
    int N, M;               // width, global height
    int nyl, j0;            // owned rows, global index of local row 0
    int jl0, jstride;       // this launch covers local rows jl0 + blockIdx.y * jstride
    int partial_base;
    double omega, homega, homegam;  // 2D:24, 365-366
    double lid6;                    // 6.0 * wd * Ulid, 2D:354-355
};

// TRT collision of one cell in the reference's operation order (2D:360-413), in place.
__host__ __device__ __forceinline__ void collide(double (&ft)[Q], const KP2 &p) {  // host: tests only
    const double drho = ft[0] + ft[1] + ft[2] + ft[3] + ft[4] + ft[5] + ft[6] + ft[7] + ft[8];  // 2D:360-361
    const double ux = ft[1] + ft[5] + ft[8] - (ft[3] + ft[6] + ft[7]);                          // 2D:362
    const double uy = ft[2] + ft[5] + ft[6] - (ft[4] + ft[7] + ft[8]);                          // 2D:363
    const double u2x15 = 1.5 * ((ux * ux) + (uy * uy));
    const double wsu2x15 = LBM2_WS * u2x15;
    const double wdu2x15 = LBM2_WD * u2x15;
    const double wsdrho = LBM2_WS * drho;
    const double wddrho = LBM2_WD * drho;
    ft[0] = ft[0] - p.omega * (ft[0] - LBM2_W0 * (drho - u2x15));  // 2D:377
#define LBM2_PAIR(a, b, cdotu, wdrho, wX, wu2)                                              \
    {                                                                                       \
        const double c_ = (cdotu);                                                          \
        const double feqTerm1 = (wdrho) + (wX) * (4.5 * (c_ * c_)) - (wu2);                 \
        const double feqTerm2 = (3.0 * (wX)) * c_;                                          \
        const double fplus = p.homega * ((ft[a] + ft[b]) - 2.0 * feqTerm1);                 \
        const double fminus = p.homegam * ((ft[a] - ft[b]) - 2.0 * feqTerm2);               \
        const double na = ft[a] - fplus - fminus;                                           \
        const double nb = ft[b] - fplus + fminus;                                           \
        ft[a] = na;                                                                         \
        ft[b] = nb;                                                                         \
    }
    LBM2_PAIR(1, 3, ux, wsdrho, LBM2_WS, wsu2x15)         // 2D:380-386
    LBM2_PAIR(2, 4, uy, wsdrho, LBM2_WS, wsu2x15)         // 2D:389-395
    LBM2_PAIR(5, 7, ux + uy, wddrho, LBM2_WD, wdu2x15)    // 2D:398-404
    LBM2_PAIR(6, 8, -ux + uy, wddrho, LBM2_WD, wdu2x15)   // 2D:407-413
#undef LBM2_PAIR
}

// grid = (ceil(N/bx), rows of this launch); block = (bx)
__global__ void __launch_bounds__(256, 4) k2_step(const __grid_constant__ KP2 p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const int jl = p.jl0 + blockIdx.y * p.jstride;
    const int jg = jl + p.j0;
    const int cell = (jl + 1) * p.N + i;
    double ft[Q];
#pragma unroll
    for (int k = 0; k < Q; k++) ft[k] = p.srcq[k][cell];
    const bool left = (i == 0), right = (i == p.N - 1), bottom = (jg == 0), top = (jg == p.M - 1);
    if (left | right | bottom | top) {
        // half-way bounce-back: the unknown populations are the cell's own opposite ones; the lid
        // rule is applied after the side walls, which decides the two top corners (2D:295-356)
        if (left)  { ft[1] = p.own[3][cell]; ft[5] = p.own[7][cell]; ft[8] = p.own[6][cell]; }
        if (right) { ft[3] = p.own[1][cell]; ft[6] = p.own[8][cell]; ft[7] = p.own[5][cell]; }
        if (bottom) { ft[2] = p.own[4][cell]; ft[5] = p.own[7][cell]; ft[6] = p.own[8][cell]; }
        if (top) {
            ft[4] = p.own[2][cell];
            ft[7] = p.own[5][cell] - p.lid6;
            ft[8] = p.own[6][cell] + p.lid6;
        }
    }
    collide(ft, p);
#pragma unroll
    for (int k = 0; k < Q; k++) p.dstq[k][cell] = ft[k];
}

// macroVars (2D:419-449): rho = 1 + sum f, u, v (momentum, no division) from the lattice `own`.
// DIFF: accumulate (du^2 + dv^2 + drho^2) against the stored fields p.u/p.v/p.rho and overwrite
// them (the reference's convergence state); !DIFF: only write the fields (output path).
template <bool DIFF>
__global__ void __launch_bounds__(256) k2_macro(const __grid_constant__ KP2 p) {
    __shared__ double wsum[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int jl = blockIdx.y;
    double d = 0.0;
    if (i < p.N) {
        const int cell = (jl + 1) * p.N + i;
        const int o = jl * p.N + i;
        double f[Q];
#pragma unroll
        for (int k = 0; k < Q; k++) f[k] = p.own[k][cell];
        const double nr = 1.0 + f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
        const double nu = f[1] - f[3] + f[5] - f[6] - f[7] + f[8];
        const double nv = f[2] - f[4] + f[5] + f[6] - f[7] - f[8];
        if (DIFF) {
            const double du = p.u[o] - nu, dv = p.v[o] - nv, dr = p.rho[o] - nr;
            d = du * du + dv * dv + dr * dr;  // pow(.,2) three times, 2D:442-443
        }
        p.u[o] = nu;
        p.rho[o] = nr;
        p.v[o] = nv;
        if (!DIFF && p.vmag) p.vmag[o] = sqrt(nu * nu + nv * nv);   // output path only
    }
    if (!DIFF) return;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; w++) s += wsum[w];
        p.partials[p.partial_base + (long long)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

}  // namespace lbm2d
