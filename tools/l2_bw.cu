// L2 -> SM bandwidth of TMA bulk loads on sm_100a: every CTA (one per SM) streams chunks of an L2-resident buffer
// into shared memory through a ring of mbarriers, nothing else.  Upper bound for the tcgen05 convolution's operand feed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/l2_bw tools/l2_bw.cu   (build outside the tree: the binary is not part of the product) && /tmp/l2_bw
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b))
                 : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(128, 1) l2_stream(const char* buf, size_t region, int chunk, int iters, int same) {
    extern __shared__ __align__(128) char smem[];
    __shared__ unsigned long long full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // same = 1: every CTA walks the SAME addresses (like the weight tiles); 0: disjoint slices (like the activations)
    const size_t span = same ? region : (region / gridDim.x) & ~(size_t)32767;      // 16-byte aligned slices, whole chunks
    const char* base = buf + (same ? 0 : (size_t)blockIdx.x * span);
    const int n_chunks = (int)(span / chunk);
    for (int it = 0; it < iters + STAGES; ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&full[s], (unsigned)((it / STAGES - 1) & 1));
        if (it < iters) {
            mbar_expect(&full[s], (unsigned)chunk);
            bulk_g2s(smem + (size_t)s * chunk, base + (size_t)(it % n_chunks) * chunk, (unsigned)chunk, &full[s]);
        }
    }
}

int main() {
    const size_t region = 64ull << 20;                       // 64 MB: resident in the 126 MB L2
    char* buf;
    cudaMalloc(&buf, region);
    cudaMemset(buf, 1, region);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000;
    for (int same = 0; same < 2; ++same)
        for (int chunk : {4096, 8192, 16384, 32768}) {
            const int stages = 6;
            const size_t smem = (size_t)stages * chunk;
            cudaFuncSetAttribute(l2_stream<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            l2_stream<6><<<sms, 128, smem>>>(buf, region, chunk, 200, same);      // warm the L2
            cudaEventRecord(e0);
            l2_stream<6><<<sms, 128, smem>>>(buf, region, chunk, iters, same);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)sms * iters * chunk;
            printf("%s addresses, %5d-byte bulk copies, 6 in flight per SM: %.3f ms  %.0f GB/s  (%.1f B/clk/SM at 1.9 GHz)  %s\n",
                   same ? "shared  " : "disjoint", chunk, ms, bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.9, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
