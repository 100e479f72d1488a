// How does HBM3e take a write stream made of R concurrent rows per CTA (rows 123 KB apart) advanced RUN bytes at a time?
// (fused update + policy kernel: a tile of 128 belief rows written 256 B per row per K block reached only 4.9 TB/s write-only while
// the same pattern reads at 6.5 TB/s — scripts/gather_stream_bench.cu, profiles/r2_gather_sweeps.md.)
// One persistent CTA per SM (optionally more), 4 warps; a CTA owns tiles of R consecutive rows and walks the columns in steps of RUN bytes:
// per step every row of the tile receives RUN contiguous bytes (st.global.v4, 512 B per warp instruction).
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a scripts/write_pattern_bench.cu -o scripts/write_pattern_bench.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128) write_kernel(float* out, long long ld, int n_tiles, int R, int run_floats, int Sp, int read_too, const float* in) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int steps = Sp / run_floats;
    float4 v = make_float4(1.f, 2.f, 3.f, (float)blockIdx.x);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int k = 0; k < steps; ++k) {
            // row segments of this step: R rows x (run_floats / 128) pieces of 512 B, dealt to the 4 warps
            const int pieces = run_floats / 128 > 0 ? run_floats / 128 : 1;
            if (run_floats >= 128) {
                for (int q = warp; q < R * pieces; q += 4) {
                    const int r = q / pieces, pc = q - r * pieces;
                    const long long off = ((long long)t * R + r) * ld + (long long)k * run_floats + pc * 128 + lane * 4;
                    if (read_too) { const float4 u = *reinterpret_cast<const float4*>(in + off); v.x += u.x; }
                    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(out + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
            } else {     // 256-byte runs: a warp instruction covers two rows
                for (int q = warp; q < R / 2; q += 4) {
                    const int r = 2 * q + (lane >> 4);
                    const long long off = ((long long)t * R + r) * ld + (long long)k * run_floats + (lane & 15) * 4;
                    if (read_too) { const float4 u = *reinterpret_cast<const float4*>(in + off); v.x += u.x; }
                    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(out + off), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
            }
        }
    }
}

int main(int argc, char** argv) {
    const int Sp = 83 * 372;
    const long long ld = argc > 1 ? atoll(argv[1]) : Sp;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long n_rows = (long long)sms * 4 * 128;
    float *out, *in;
    CK(cudaMalloc(&out, n_rows * ld * 4));
    CK(cudaMalloc(&in, n_rows * ld * 4));
    CK(cudaMemset(in, 0, n_rows * ld * 4));
    const int shapes[][2] = {{128, 64}, {64, 128}, {32, 256}, {16, 512}, {8, 1024}, {4, 2048}, {2, 4096}, {1, 7680}};
    for (int ctas = 1; ctas <= 4; ctas *= 2)
    for (int read_too = 0; read_too < 2; ++read_too)
    for (auto& s : shapes) {
        const int R = s[0], run = s[1];
        const int n_tiles = (int)(n_rows / R);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        write_kernel<<<sms * ctas, 128>>>(out, ld, n_tiles, R, run, Sp, read_too, in);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 3; ++i) write_kernel<<<sms * ctas, 128>>>(out, ld, n_tiles, R, run, Sp, read_too, in);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 3;
        const double bytes = (double)n_rows * (Sp / run) * run * 4 * (read_too ? 2 : 1);
        printf("ld %lld ctas/SM %d %s rows/CTA %4d run %5d B: %7.3f ms %7.1f GB/s\n", ld, ctas, read_too ? "copy " : "write", R, run * 4, ms, bytes / ms / 1e6);
    }
    return 0;
}
