// tools/tma_probe.cu -- stand-alone probe of the TMA tensor copy used by k_step_pair (STAGE 4):
// 4-D fp64 tensor (x, y, z, q), box (32, TY, 1, 1) at arbitrary (also negative / odd) coordinates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe.bin tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BX, int TY>
__global__ void probe(const __grid_constant__ CUtensorMap tm, double *out, int x, int y, int z, int q) {
    extern __shared__ __align__(128) double sm[];
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(sm + BX * TY);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(BX * TY * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&tm), "r"(mbar), "r"(x), "r"(y), "r"(z), "r"(q) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(mbar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BX * TY; i += blockDim.x) out[i] = sm[i];
}

template <int BX, int TY>
static int run(encode_fn enc, int N, int nz, int x, int y, int z, int q) {
    const size_t n = (size_t)N * N * nz * 19;
    double *h = (double *)malloc(n * 8), *d, *o, ho[BX * TY];
    for (size_t i = 0; i < n; i++) h[i] = (double)i + 0.5;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, BX * TY * 8);
    cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)nz, 19};
    const cuuint64_t strides[3] = {(cuuint64_t)N * 8, (cuuint64_t)N * N * 8, (cuuint64_t)N * N * nz * 8};
    const cuuint32_t box[4] = {BX, TY, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("BX=%d TY=%d N=%d: encode failed %d\n", BX, TY, N, (int)r); return 1; }
    const int smem = BX * TY * 8 + 16;
    cudaFuncSetAttribute(probe<BX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<BX, TY><<<1, 128, smem>>>(tm, o, x, y, z, q);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("BX=%d TY=%d N=%d at (%d,%d,%d,%d): %s\n", BX, TY, N, x, y, z, q, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < TY; j++)
        for (int i = 0; i < BX; i++) {
            const int gx = x + i, gy = y + j;
            double want = 0.0;
            if (gx >= 0 && gx < N && gy >= 0 && gy < N && z >= 0 && z < nz) want = h[(((size_t)q * nz + z) * N + gy) * N + gx];
            if (ho[j * BX + i] != want) bad++;
        }
    printf("BX=%d TY=%d N=%d at (%d,%d,%d,%d): %s (%d mismatches)\n", BX, TY, N, x, y, z, q, bad ? "WRONG" : "ok", bad);
    cudaFree(d); cudaFree(o); free(h);
    return bad ? 3 : 0;
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no encoder\n"); return 1; }
    encode_fn enc = (encode_fn)fn;
    int rc = 0;
    // FINDING (B200, driver 580): the box must start on a 16-byte boundary -- an odd x coordinate of an
    // fp64 tensor raises "illegal instruction" (sticky), so the odd-start cases come last.
    rc |= run<16, 4>(enc, 64, 6, 0, 0, 1, 3);      // small aligned box
    rc |= run<16, 4>(enc, 64, 6, 2, 1, 1, 3);      // 16-byte aligned start
    rc |= run<34, 16>(enc, 64, 6, -2, -1, 0, 0);   // the kernel's box: 34 wide, negative even start
    rc |= run<34, 16>(enc, 64, 6, 28, -2, 2, 18);  // the kernel's box, even start, rows out of range
    rc |= run<34, 16>(enc, 12, 14, -2, -1, 1, 5);  // box larger than the tensor
    rc |= run<34, 10>(enc, 512, 4, 478, 503, 3, 7);
    rc |= run<16, 4>(enc, 64, 6, 3, 1, 1, 3);      // odd start (8-byte aligned only): expected to fail
    return rc;
}
