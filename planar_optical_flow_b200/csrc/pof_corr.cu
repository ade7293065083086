// Windowed patch correlation of the scan-pair flow prototype.
//
// Replaces Prototype._fusion (/root/reference/src/depracted/model/prototype.py:118-156): the reference
// gathers K-tap patches of both feature maps ([B, C*K, N]), multiplies them into a dense [N, N]
// correlation matrix and then keeps the 2D+1 entries per point within `max_displacement` — the same
// dense-then-window pattern as the DR-SPAAM gate.  With the reference's index clamping
//
//     out[b, d + D, i] = sum_c sum_{k=-h..h} f1[b, c, clamp(i + k)] * f2[b, c, clamp(clamp(i + d) + k)]
//
// Forward: one warp per point, one lane per displacement (2D+1 <= 32), channels walked serially, so no
// reduction and no [N, N] matrix.  Backward: one CTA per (sample, channel chunk), one thread per channel;
// each thread keeps its channel's two feature rows and two gradient rows in shared memory and replays the
// (i, d, k) triples, so the clamped scatter needs no atomics and the result is deterministic.
#include "pof_common.cuh"

namespace pof {
namespace {

__device__ __forceinline__ int clampi(int v, int hi) { return min(max(v, 0), hi); }

__global__ void __launch_bounds__(256) patch_corr_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int B, int C,
                                                             int N, int h, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= (long long)B * N) return;
    const int b = (int)(warp / N), i = (int)(warp - (long long)b * N);
    const int W = 2 * D + 1;
    const int d = min(lane, W - 1) - D;                        // lanes beyond the window repeat the last one, unused
    const int j = clampi(i + d, N - 1);
    const float* r1 = f1 + (size_t)b * C * N;
    const float* r2 = f2 + (size_t)b * C * N;
    float acc = 0.f;
    for (int c = 0; c < C; ++c, r1 += N, r2 += N)
        for (int k = -h; k <= h; ++k)
            acc = fmaf(__ldg(r1 + clampi(i + k, N - 1)), __ldg(r2 + clampi(j + k, N - 1)), acc);
    if (lane < W) out[((size_t)b * W + lane) * N + i] = acc;
}

// grid (B, ceil(C / CC)), CC threads; dynamic shared memory: g [W][N] | f1 | f2 | a1 | a2, each [CC][pitch]
__global__ void patch_corr_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ g_out, int C,
                                      int N, int h, int D, int pitch, float* __restrict__ g1, float* __restrict__ g2) {
    extern __shared__ float sm[];
    const int W = 2 * D + 1, CC = blockDim.x;
    float* g = sm;
    float* s1 = g + W * N;
    float* s2 = s1 + CC * pitch;
    float* a1 = s2 + CC * pitch;
    float* a2 = a1 + CC * pitch;
    const int b = blockIdx.x, c0 = blockIdx.y * CC, cc = min(CC, C - c0);
    for (int t = threadIdx.x; t < W * N; t += CC) g[t] = __ldg(g_out + (size_t)b * W * N + t);
    const size_t base = ((size_t)b * C + c0) * N;
    for (int t = threadIdx.x; t < cc * N; t += CC) {           // coalesced: the chunk is contiguous in global memory
        const int c = t / N, p = t - c * N;
        s1[c * pitch + p] = __ldg(f1 + base + t);
        s2[c * pitch + p] = __ldg(f2 + base + t);
        a1[c * pitch + p] = 0.f;
        a2[c * pitch + p] = 0.f;
    }
    __syncthreads();
    if ((int)threadIdx.x < cc) {
        const float* x1 = s1 + threadIdx.x * pitch;
        const float* x2 = s2 + threadIdx.x * pitch;
        float* y1 = a1 + threadIdx.x * pitch;
        float* y2 = a2 + threadIdx.x * pitch;
        for (int i = 0; i < N; ++i)
            for (int d = -D; d <= D; ++d) {
                const float gv = g[(d + D) * N + i];
                const int j = clampi(i + d, N - 1);
                for (int k = -h; k <= h; ++k) {
                    const int p1 = clampi(i + k, N - 1), p2 = clampi(j + k, N - 1);
                    y1[p1] = fmaf(gv, x2[p2], y1[p1]);
                    y2[p2] = fmaf(gv, x1[p1], y2[p2]);
                }
            }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cc * N; t += CC) {
        const int c = t / N, p = t - c * N;
        g1[base + t] = a1[c * pitch + p];
        g2[base + t] = a2[c * pitch + p];
    }
}

int bwd_chunk(int C, int N, int W, int pitch, size_t* smem) {
    for (int cc = 256; cc >= 32; cc -= 32) {
        const size_t need = ((size_t)W * N + (size_t)4 * cc * pitch) * sizeof(float);
        if (need <= 200 * 1024) {
            *smem = need;
            return cc < C ? cc : ((C + 31) / 32) * 32;
        }
    }
    return 0;
}

}  // namespace
}  // namespace pof

extern "C" {

int pof_patch_corr_fwd(const float* feat1, const float* feat2, int B, int C, int N, int kernel_size, int max_displacement, float* out,
                       void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(feat1 && feat2 && out, POF_ERR_NULL_POINTER, "pof_patch_corr_fwd: null pointer");
    POF_REQUIRE(B > 0 && C >= 1 && N >= 1, POF_ERR_BAD_SHAPE, "pof_patch_corr_fwd: bad shape B=%d C=%d N=%d", B, C, N);
    POF_REQUIRE(kernel_size >= 1 && (kernel_size & 1) && max_displacement >= 0 && 2 * max_displacement + 1 <= 32, POF_ERR_BAD_PARAM,
                "pof_patch_corr_fwd: kernel_size must be odd and 2*max_displacement+1 <= 32 (got %d, %d)", kernel_size, max_displacement);
    const long long warps = (long long)B * N;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    patch_corr_fwd_kernel<<<grid, 256, 0, stream>>>(feat1, feat2, B, C, N, kernel_size / 2, max_displacement, out);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_patch_corr_bwd(const float* feat1, const float* feat2, const float* grad_out, int B, int C, int N, int kernel_size,
                       int max_displacement, float* grad_feat1, float* grad_feat2, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(feat1 && feat2 && grad_out && grad_feat1 && grad_feat2, POF_ERR_NULL_POINTER, "pof_patch_corr_bwd: null pointer");
    POF_REQUIRE(B > 0 && C >= 1 && N >= 1, POF_ERR_BAD_SHAPE, "pof_patch_corr_bwd: bad shape B=%d C=%d N=%d", B, C, N);
    POF_REQUIRE(kernel_size >= 1 && (kernel_size & 1) && max_displacement >= 0 && 2 * max_displacement + 1 <= 32, POF_ERR_BAD_PARAM,
                "pof_patch_corr_bwd: kernel_size must be odd and 2*max_displacement+1 <= 32 (got %d, %d)", kernel_size, max_displacement);
    const int W = 2 * max_displacement + 1, pitch = N | 1;          // odd pitch: one thread per row without bank conflicts
    size_t smem = 0;
    const int cc = bwd_chunk(C, N, W, pitch, &smem);
    POF_REQUIRE(cc > 0, POF_ERR_UNSUPPORTED, "pof_patch_corr_bwd: %d points per feature row do not fit shared memory", N);
    smem = ((size_t)W * N + (size_t)4 * cc * pitch) * sizeof(float);
    static bool attr_set[64] = {false};
    int dev = 0;
    POF_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        POF_CUDA(cudaFuncSetAttribute(patch_corr_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[dev] = true;
    }
    patch_corr_bwd_kernel<<<dim3((unsigned)B, (unsigned)((C + cc - 1) / cc)), cc, smem, stream>>>(
        feat1, feat2, grad_out, C, N, kernel_size / 2, max_displacement, pitch, grad_feat1, grad_feat2);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
