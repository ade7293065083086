// Store-pattern microbenchmark: how fast can the SMs push 16-byte stores to L2/HBM when a warp instruction covers
// SEG-byte contiguous segments of rows that are PITCH bytes apart (the convolution epilogue writes 64-byte segments).
//   nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/store_pattern_bench tools/store_pattern_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SEG>   // bytes per row segment written by one warp instruction: 16, 64, 128, 512
__global__ void store_kernel(char* out, long long rows, int pitch, int iters) {
    constexpr int lanes_per_row = SEG / 16, rows_per_instr = 32 / lanes_per_row;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
    // a warp owns 32 consecutive rows and walks along them in SEG-byte steps, like an epilogue warp walking over its columns
    for (long long r0 = warp * 32; r0 + 32 <= rows; r0 += n_warps * 32)
        for (int c = 0; c < pitch; c += SEG)
            for (int rr = 0; rr < 32; rr += rows_per_instr) {
                char* p = out + (r0 + rr + lane / lanes_per_row) * pitch + c + (lane % lanes_per_row) * 16;
                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
}

template <int SEG>
void run(char* buf, long long rows, int pitch, int ctas_per_sm, int threads) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) store_kernel<SEG><<<sms * ctas_per_sm, threads>>>(buf, rows, pitch, 1);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) store_kernel<SEG><<<sms * ctas_per_sm, threads>>>(buf, rows, pitch, 1);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 10;
    const double bytes = (double)(rows / 32 * 32) * pitch;
    printf("segment %3d B, row pitch %4d B, %d x %d threads per SM: %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)\n", SEG, pitch, ctas_per_sm,
           threads, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.9);
}

int main() {
    const long long bytes = 1ll << 30;
    char* buf = nullptr;
    cudaMalloc(&buf, bytes);
    for (int pitch : {256, 512, 1024})
        for (int warps : {8, 32}) {
            const long long rows = bytes / pitch;
            run<16>(buf, rows, pitch, 1, warps * 32);
            run<64>(buf, rows, pitch, 1, warps * 32);
            run<128>(buf, rows, pitch, 1, warps * 32);
            if (pitch >= 512) run<512>(buf, rows, pitch, 1, warps * 32);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
