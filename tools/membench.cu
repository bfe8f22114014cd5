// tools/membench.cu -- measurement aid (not product): the D3Q19 access pattern (19 shifted loads,
// 19 aligned stores per cell, fp64) with a tunable number of fp64 operations per cell, to see how
// the sustained (power-capped) bandwidth of a B200 depends on the arithmetic carried by the stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/membench.bin tools/membench.cu
//   tools/membench.bin [N=256] [seconds=3]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct P { const double *s[19]; double *d[19]; int N; long long N2; };

template <int OPS>   // OPS = extra fp64 add/mul pairs per population
__global__ void __launch_bounds__(128, 4) k(const __grid_constant__ P p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int cell = (blockIdx.z + 1) * (int)p.N2 + blockIdx.y * p.N + i;
    double f[19];
#pragma unroll
    for (int q = 0; q < 19; q++) f[q] = p.s[q][cell];
    if (OPS > 0) {
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < 19; q++) a += f[q];
#pragma unroll
        for (int q = 0; q < 19; q++) {
            double x = f[q];
#pragma unroll
            for (int o = 0; o < OPS; o++) { x = x * 0.5 + a * 0.02; }
            f[q] = x;
        }
    }
#pragma unroll
    for (int q = 0; q < 19; q++) p.d[q][cell] = f[q];
}

__global__ void fill_random(double *a, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    unsigned long long x = (unsigned long long)t * 6364136223846793005ULL + 1442695040888963407ULL;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
    a[t] = 0.01 + 0.04 * (double)(x >> 11) * (1.0 / 9007199254740992.0);   // random mantissas, LBM-like magnitudes
}

// 128-bit variant: two cells per thread, double2 loads/stores on the populations whose x-shift is 0
// (9 of 19), two 8-byte loads on the x-shifted ones (10 of 19), double2 stores for all 19.
template <int OPS>
__global__ void __launch_bounds__(128, 2) k2(const __grid_constant__ P p) {
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int cell = (blockIdx.z + 1) * (int)p.N2 + blockIdx.y * p.N + i;
    double f[19][2];
#pragma unroll
    for (int q = 0; q < 19; q++) {
        const bool xs = (q == 1 || q == 2 || (q >= 7 && q <= 10) || (q >= 13 && q <= 16));
        if (xs) { f[q][0] = p.s[q][cell]; f[q][1] = p.s[q][cell + 1]; }
        else { const double2 v = *reinterpret_cast<const double2 *>(p.s[q] + cell); f[q][0] = v.x; f[q][1] = v.y; }
    }
    if (OPS > 0) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            double a = 0.0;
#pragma unroll
            for (int q = 0; q < 19; q++) a += f[q][c];
#pragma unroll
            for (int q = 0; q < 19; q++) {
                double x = f[q][c];
#pragma unroll
                for (int o = 0; o < OPS; o++) { x = x * 0.5 + a * 0.02; }
                f[q][c] = x;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 19; q++) *reinterpret_cast<double2 *>(p.d[q] + cell) = make_double2(f[q][0], f[q][1]);
}

#define LAUNCH_IMPL(OPS, V2, grid, blk, arg) do { if (V2) k2<OPS><<<grid, blk>>>(arg); else k<OPS><<<grid, blk>>>(arg); } while (0)
template <int OPS, bool V2 = false>
static void run(const P &pa, const P &pb, int N, double seconds) {
    dim3 grid(V2 ? N / 256 : N / 128, N, N), blk(128);

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 20; w++) { LAUNCH_IMPL(OPS, V2, grid, blk, pa); LAUNCH_IMPL(OPS, V2, grid, blk, pb); }
    cudaDeviceSynchronize();
    // burst: first 100 launches; sustained: keep going for `seconds`, report the last 200 launches
    float ms_first = 0, ms_last = 0;
    cudaEventRecord(e0);
    for (int w = 0; w < 50; w++) { LAUNCH_IMPL(OPS, V2, grid, blk, pa); LAUNCH_IMPL(OPS, V2, grid, blk, pb); }
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_first, e0, e1);
    double per = ms_first / 100.0;
    int nl = (int)(seconds * 1e3 / per / 2);
    for (int w = 0; w < nl; w++) { LAUNCH_IMPL(OPS, V2, grid, blk, pa); LAUNCH_IMPL(OPS, V2, grid, blk, pb); }
    cudaEventRecord(e0);
    for (int w = 0; w < 100; w++) { LAUNCH_IMPL(OPS, V2, grid, blk, pa); LAUNCH_IMPL(OPS, V2, grid, blk, pb); }
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_last, e0, e1);
    double bytes = 304.0 * N * N * (double)N;
    printf("{\"vec128\": %d, \"fp64_ops_per_cell\": %d, \"burst_GBs\": %.1f, \"sustained_GBs\": %.1f, \"burst_ms\": %.4f, \"sustained_ms\": %.4f}\n",
           (int)V2, OPS * 38 + (OPS ? 19 : 0), bytes / (ms_first / 100 * 1e-3) / 1e9, bytes / (ms_last / 200 * 1e-3) / 1e9, ms_first / 100, ms_last / 200);
    fflush(stdout);
}

int main(int argc, char **argv) {
    int N = argc > 1 ? atoi(argv[1]) : 256;
    double sec = argc > 2 ? atof(argv[2]) : 3.0;
    long long N2 = (long long)N * N, PS = (long long)(N + 2) * N2, pad = N2 + N + 16;
    double *A, *B;
    cudaMalloc(&A, (19 * PS + 2 * pad) * 8); cudaMalloc(&B, (19 * PS + 2 * pad) * 8);
    {
        long long n = 19 * PS + 2 * pad;
        fill_random<<<(unsigned)((n + 255) / 256), 256>>>(A, n);
        fill_random<<<(unsigned)((n + 255) / 256), 256>>>(B, n);
        cudaDeviceSynchronize();
    }
    static const int cx[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0};
    static const int cy[19] = {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1};
    static const int cz[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1};
    P pa, pb;
    pa.N = pb.N = N; pa.N2 = pb.N2 = N2;
    for (int q = 0; q < 19; q++) {
        long long sh = cx[q] + (long long)cy[q] * N + (long long)cz[q] * N2;
        pa.s[q] = A + pad + q * PS - sh; pa.d[q] = B + pad + q * PS;
        pb.s[q] = B + pad + q * PS - sh; pb.d[q] = A + pad + q * PS;
    }
    run<0>(pa, pb, N, sec);
    run<2>(pa, pb, N, sec);
    run<4>(pa, pb, N, sec);
    run<8>(pa, pb, N, sec);
    run<12>(pa, pb, N, sec);
    run<0, true>(pa, pb, N, sec);
    run<8, true>(pa, pb, N, sec);
    return 0;
}
