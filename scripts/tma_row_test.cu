// Micro-test: single-row TMA tensor copies (box = 128 floats x 1 row, no swizzle) with arbitrary (negative / overhanging)
// start columns: does the copy complete, are out-of-range columns zero, is the full box counted on the mbarrier?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int col, int row, float* out, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    float* buf = reinterpret_cast<float*>(smem + (base - smem_u32(smem)));
    const uint32_t bar = base + 4096;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = -7.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(512) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(base + (variant ? 512u : 0u)), "l"(&tm), "r"(col), "r"(row), "r"(bar) : "memory");
    }
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
        if (clock64() - t0 > 200000000LL) { if (threadIdx.x == 0) printf("TIMEOUT col %d\n", col); break; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int Sp = 360, rows = 300; const long long ld = Sp;
    float* A; CK(cudaMalloc(&A, rows * ld * 4));
    float* h = (float*)malloc(rows * ld * 4);
    for (int i = 0; i < rows * ld; ++i) h[i] = (float)(i % 1000) + 1.f;
    CK(cudaMemcpy(A, h, rows * ld * 4, cudaMemcpyHostToDevice));
    float* out; CK(cudaMalloc(&out, 1024));
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    for (unsigned long long outer : {300ull, 1ull << 20, 1ull << 31}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)Sp, outer}; cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {128, 1}; cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("outer %llu encode rc %d\n", outer, (int)r);
        if (r != CUDA_SUCCESS) continue;
        for (int col : {0, 128, -24, 280, 356, 4, 1}) {
            k<<<1, 128, 8192>>>(tm, col, 5, out, 0);
            cudaError_t e = cudaDeviceSynchronize();
            float ho[256]; cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int i = 0; i < 128; ++i) { int c = col + i; float want = (c >= 0 && c < Sp) ? h[5 * ld + c] : 0.f; if (ho[i] != want) ++bad; }
            printf("  col %4d: %s, mismatches %d, first %g %g %g, beyond-box %g\n", col, cudaGetErrorString(e), bad, ho[0], ho[1], ho[2], ho[128]);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
