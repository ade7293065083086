// Training-mode BatchNorm + LeakyReLU (+ max-pool over row pairs) of the DR-SPAAM conv layers, forward and backward,
// on channels-last activations [rows = cutout x position, C].
//
// The reference's training step (/root/reference/src/depracted/model/dr_spaam.py:8-12 `_conv` = Conv1d + BatchNorm1d +
// LeakyReLU(0.1), :81-97 max_pool1d(2) after each block; bin/train_dr_spaam.py) runs, per layer, cuDNN's batch-norm kernel,
// an in-place LeakyReLU and PyTorch's max-pool kernel forward, and the three matching kernels backward: 6.5 and ~9 passes
// over the layer's activations.  The launch list of a step (profiles/r2_train_launch_summary.txt) shows them at 24 % + 9 % +
// 18 % of the step, against ~35 % for the TF32 GEMMs they sit between.  Here the three are one operator:
//
//   forward   pof_bn_act_stats   one read of y: per-channel sum and sum of squares (fp32 partial sums per thread over <= 64
//                                rows, fp64 atomics per CTA)
//             pof_bn_act_fwd     one read of y, one write of z (half the rows when pooled):
//                                z = max over the pair of lrelu((y - mean) * invstd * gamma + beta); running statistics updated
//   backward  pof_bn_act_bwd_reduce   reads y and dz: recomputes the pre-activation, its LeakyReLU slope and the pair's arg-max
//                                (first on ties, like max_pool's saved indices) -> d_beta = sum d_hat, d_gamma = sum d_hat * x_hat
//             pof_bn_act_bwd     reads y and dz again, writes dx = gamma * invstd * (d_hat - d_beta/n - x_hat * d_gamma/n)
//
// Nothing but y, mean and invstd is saved for the backward (PyTorch saves the conv output, the batch-norm output, the pool
// input and its indices).  All four are streaming kernels: a thread owns 4 adjacent channels and walks rows; 128-bit accesses.
#include "pof_common.cuh"

namespace pof {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnRowsPerThread = 16;          // rows a thread reduces in fp32 before the partial sums go to fp64 (64 rows per thread left the
                                              // reductions at 197 CTAs = 1.3 waves, latency bound at a third of the copy bandwidth)

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- forward statistics -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ y, long long rows /* per group */, int C,
                                                              double* __restrict__ sums /* [G][2][C] */) {
    extern __shared__ float red[];                         // [2][rpb][C]
    y += (size_t)blockIdx.y * rows * C;                    // group g = blockIdx.y: its own rows, its own statistics
    sums += (size_t)blockIdx.y * 2 * C;
    const int c4n = C >> 2, rpb = kBnThreads / c4n;
    const int c = (threadIdx.x % c4n) << 2, rg = threadIdx.x / c4n;
    const long long slab = (long long)rpb * kBnRowsPerThread;
    for (long long r0 = (long long)blockIdx.x * slab; r0 < rows; r0 += (long long)gridDim.x * slab) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
        if (rg < rpb) {
#pragma unroll 8
            for (int k = 0; k < kBnRowsPerThread; ++k) {
                const long long r = r0 + (long long)k * rpb + rg;
                if (r < rows) {
                    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(y + r * C + c));
                    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
                }
            }
            *reinterpret_cast<float4*>(red + (size_t)rg * C + c) = s;
            *reinterpret_cast<float4*>(red + (size_t)(rpb + rg) * C + c) = q;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * C; t += kBnThreads) {       // t < C: sums, else sums of squares
            const int which = t / C, ch = t - which * C;
            double acc = 0.0;
            for (int g = 0; g < rpb; ++g) acc += (double)red[(size_t)(which * rpb + g) * C + ch];
            atomicAdd(sums + t, acc);
        }
        __syncthreads();
    }
}

struct BnArgs {
    const float* y;
    const float* dz;
    const double* sums;      // forward: [sum | sum of squares]; backward: [sum d_hat | sum d_hat * x_hat]
    const float* gamma;
    const float* beta;
    float* mean;             // [C] written by the forward, read by the backward
    float* invstd;           // [C]
    float* running_mean;     // [C] or null
    float* running_var;      // [C] or null
    float* out;              // forward: z [rows / pool, C]; backward: dx [rows, C]
    float* dgamma;
    float* dbeta;
    long long rows;          // rows PER GROUP
    int C, pool, groups;     // groups: consecutive blocks of `rows` rows, each normalised with its own batch statistics
    float eps, slope, momentum;
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// mean / invstd / scale / shift of this thread's four channels from the reduced sums (every thread recomputes them: 8 loads)
__device__ __forceinline__ void channel_stats(const BnArgs& a, int g, int c, float (&mean)[4], float (&invstd)[4]) {
    const double n = (double)a.rows;
    const double* sums = a.sums + (size_t)g * 2 * a.C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double m = sums[c + j] / n;
        double var = sums[a.C + c + j] / n - m * m;         // biased variance, as batch_norm normalises with
        if (var < 0.0) var = 0.0;
        mean[j] = (float)m;
        invstd[j] = (float)(1.0 / sqrt(var + (double)a.eps));
    }
}

template <int POOL>
__global__ void __launch_bounds__(kBnThreads) bn_act_fwd_kernel(const BnArgs a) {
    const int C = a.C, c4n = C >> 2, rpb = kBnThreads / c4n;
    const int c = (threadIdx.x % c4n) << 2, rg = threadIdx.x / c4n;
    const int grp = blockIdx.y;
    float mean[4], invstd[4], scale[4], shift[4];
    channel_stats(a, grp, c, mean, invstd);
    const float4 g = ld4(a.gamma + c), b = ld4(a.beta + c);
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { scale[j] = invstd[j] * gg[j]; shift[j] = bb[j]; }
    if (blockIdx.x == 0 && rg == 0) {                       // one writer per (group, channel): the saved statistics
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a.mean[(size_t)grp * C + c + j] = mean[j];
            a.invstd[(size_t)grp * C + c + j] = invstd[j];
        }
        if (grp == 0 && a.running_mean) {                   // running statistics: one update per group, IN GROUP ORDER (= the order of the
            const double n = (double)a.rows;                // reference's separate batch_norm calls, one per scan)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float rm = a.running_mean[c + j], rv = a.running_var[c + j];
                for (int q = 0; q < a.groups; ++q) {
                    const double* sq = a.sums + (size_t)q * 2 * C;
                    const double m = sq[c + j] / n;
                    double var = sq[C + c + j] / n - m * m;
                    if (var < 0.0) var = 0.0;
                    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
                    rm = (1.f - a.momentum) * rm + a.momentum * (float)m;
                    rv = (1.f - a.momentum) * rv + a.momentum * (float)unbiased;
                }
                a.running_mean[c + j] = rm;
                a.running_var[c + j] = rv;
            }
        }
    }
    const long long rows_out = a.rows / POOL;
    if (rg >= rpb) return;
    const float* y = a.y + (size_t)grp * a.rows * C;
    float* out = a.out + (size_t)grp * rows_out * C;
    for (long long r = (long long)blockIdx.x * rpb + rg; r < rows_out; r += (long long)gridDim.x * rpb) {
        const float4 v0 = ld_stream_f4(reinterpret_cast<const float4*>(y + r * POOL * C + c));
        const float x0[4] = {v0.x, v0.y, v0.z, v0.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = lrelu(fmaf(x0[j] - mean[j], scale[j], shift[j]), a.slope);
        if (POOL == 2) {
            const float4 v1 = ld_stream_f4(reinterpret_cast<const float4*>(y + (r * POOL + 1) * C + c));
            const float x1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], lrelu(fmaf(x1[j] - mean[j], scale[j], shift[j]), a.slope));
        }
        *reinterpret_cast<float4*>(out + r * C + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// Gradient with respect to the batch-norm output of the POOL rows behind pooled row r, channel quad c:
// d_hat[p][j], and the normalised inputs x_hat[p][j].
template <int POOL>
__device__ __forceinline__ void pair_grads(const BnArgs& a, const float* __restrict__ y, const float* __restrict__ dz, long long r, int c,
                                           const float (&mean)[4], const float (&invstd)[4],
                                           const float (&gg)[4], const float (&bb)[4], float (&dh)[POOL][4], float (&xh)[POOL][4]) {
    const float4 d4 = ld_stream_f4(reinterpret_cast<const float4*>(dz + r * a.C + c));
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    float act[POOL][4];
#pragma unroll
    for (int p = 0; p < POOL; ++p) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(y + (r * POOL + p) * a.C + c));
        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xh[p][j] = (x[j] - mean[j]) * invstd[j];
            const float pre = fmaf(x[j] - mean[j], invstd[j] * gg[j], bb[j]);      // the forward's expression, bit for bit
            act[p][j] = lrelu(pre, a.slope);
            dh[p][j] = d[j] * (pre > 0.f ? 1.f : a.slope);
        }
    }
    if (POOL == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool second = act[1][j] > act[0][j];       // first wins ties, like max_pool's saved indices
            dh[second ? 0 : 1][j] = 0.f;
        }
    }
}

template <int POOL>
__global__ void __launch_bounds__(kBnThreads) bn_act_bwd_reduce_kernel(const BnArgs a, double* __restrict__ sums /* [2][C] */) {
    extern __shared__ float red[];                         // [2][rpb][C]
    const int C = a.C, c4n = C >> 2, rpb = kBnThreads / c4n;
    const int c = (threadIdx.x % c4n) << 2, rg = threadIdx.x / c4n;
    const int grp = blockIdx.y;
    float mean[4], invstd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mean[j] = __ldg(a.mean + (size_t)grp * C + c + j); invstd[j] = __ldg(a.invstd + (size_t)grp * C + c + j); }
    const float4 g = ld4(a.gamma + c), b = ld4(a.beta + c);
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
    const long long rows_out = a.rows / POOL;
    const float* y = a.y + (size_t)grp * a.rows * C;
    const float* dz = a.dz + (size_t)grp * rows_out * C;
    sums += (size_t)grp * 2 * C;
    const long long slab = (long long)rpb * (kBnRowsPerThread / POOL);
    for (long long r0 = (long long)blockIdx.x * slab; r0 < rows_out; r0 += (long long)gridDim.x * slab) {
        float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
        if (rg < rpb) {
            for (int k = 0; k < kBnRowsPerThread / POOL; ++k) {
                const long long r = r0 + (long long)k * rpb + rg;
                if (r < rows_out) {
                    float dh[POOL][4], xh[POOL][4];
                    pair_grads<POOL>(a, y, dz, r, c, mean, invstd, gg, bb, dh, xh);
#pragma unroll
                    for (int p = 0; p < POOL; ++p)
#pragma unroll
                        for (int j = 0; j < 4; ++j) { s[j] += dh[p][j]; q[j] = fmaf(dh[p][j], xh[p][j], q[j]); }
                }
            }
            *reinterpret_cast<float4*>(red + (size_t)rg * C + c) = make_float4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<float4*>(red + (size_t)(rpb + rg) * C + c) = make_float4(q[0], q[1], q[2], q[3]);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * C; t += kBnThreads) {
            const int which = t / C, ch = t - which * C;
            double acc = 0.0;
            for (int gi = 0; gi < rpb; ++gi) acc += (double)red[(size_t)(which * rpb + gi) * C + ch];
            atomicAdd(sums + t, acc);
        }
        __syncthreads();
    }
}

template <int POOL>
__global__ void __launch_bounds__(kBnThreads) bn_act_bwd_kernel(const BnArgs a) {
    const int C = a.C, c4n = C >> 2, rpb = kBnThreads / c4n;
    const int c = (threadIdx.x % c4n) << 2, rg = threadIdx.x / c4n;
    const int grp = blockIdx.y;
    float mean[4], invstd[4], db[4], dg[4];
    const double n = (double)a.rows;
    const double* sums = a.sums + (size_t)grp * 2 * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mean[j] = __ldg(a.mean + (size_t)grp * C + c + j);
        invstd[j] = __ldg(a.invstd + (size_t)grp * C + c + j);
        db[j] = (float)(sums[c + j] / n);
        dg[j] = (float)(sums[C + c + j] / n);
    }
    const float4 g = ld4(a.gamma + c), b = ld4(a.beta + c);
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
    if (blockIdx.x == 0 && grp == 0 && rg == 0) {           // gamma and beta are shared by the groups: their gradients add up
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double tb = 0.0, tg = 0.0;
            for (int q = 0; q < a.groups; ++q) { tb += a.sums[(size_t)q * 2 * C + c + j]; tg += a.sums[(size_t)q * 2 * C + C + c + j]; }
            a.dbeta[c + j] = (float)tb;
            a.dgamma[c + j] = (float)tg;
        }
    }
    const long long rows_out = a.rows / POOL;
    if (rg >= rpb) return;
    const float* y = a.y + (size_t)grp * a.rows * C;
    const float* dz = a.dz + (size_t)grp * rows_out * C;
    float* out = a.out + (size_t)grp * a.rows * C;
    for (long long r = (long long)blockIdx.x * rpb + rg; r < rows_out; r += (long long)gridDim.x * rpb) {
        float dh[POOL][4], xh[POOL][4];
        pair_grads<POOL>(a, y, dz, r, c, mean, invstd, gg, bb, dh, xh);
#pragma unroll
        for (int p = 0; p < POOL; ++p) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = gg[j] * invstd[j] * (dh[p][j] - db[j] - xh[p][j] * dg[j]);
            *reinterpret_cast<float4*>(out + (r * POOL + p) * C + c) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---- weight gradient of the first layer (Conv1d(1 -> C, k = 3, p = 1), dr_spaam.py:49) ----------------------------
// dW[c][k] = sum over rows (m, l) of dy[m, l, c] * x[m, l + k - 1]  (zero outside the cutout).  cuDNN has no tensor-core
// engine for one input channel and takes 0.28 ms per call (eleven calls per training step); this is one read of dy.
__global__ void __launch_bounds__(kBnThreads) conv_first_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                     long long rows, int P, int C, double* __restrict__ sums /* [3][C] */) {
    extern __shared__ float red[];                         // [3][rpb][C]
    const int c4n = C >> 2, rpb = kBnThreads / c4n;
    const int c = (threadIdx.x % c4n) << 2, rg = threadIdx.x / c4n;
    const long long slab = (long long)rpb * kBnRowsPerThread;
    for (long long r0 = (long long)blockIdx.x * slab; r0 < rows; r0 += (long long)gridDim.x * slab) {
        float acc[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
        for (int k = 0; k < kBnRowsPerThread; ++k) {
            const long long r = r0 + (long long)k * rpb + rg;
            if (r < rows) {
                const int l = (int)(r % P);
                const float xc = __ldg(x + r);
                const float xl = l > 0 ? __ldg(x + r - 1) : 0.f;
                const float xr = l < P - 1 ? __ldg(x + r + 1) : 0.f;
                const float4 g = ld_stream_f4(reinterpret_cast<const float4*>(dy + r * C + c));
                const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[0][j] = fmaf(gv[j], xl, acc[0][j]);
                    acc[1][j] = fmaf(gv[j], xc, acc[1][j]);
                    acc[2][j] = fmaf(gv[j], xr, acc[2][j]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t)
            *reinterpret_cast<float4*>(red + (size_t)(t * rpb + rg) * C + c) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        __syncthreads();
        for (int t = threadIdx.x; t < 3 * C; t += kBnThreads) {
            const int tap = t / C, ch = t - tap * C;
            double a2 = 0.0;
            for (int g = 0; g < rpb; ++g) a2 += (double)red[(size_t)(tap * rpb + g) * C + ch];
            atomicAdd(sums + t, a2);
        }
        __syncthreads();
    }
}
__global__ void conv_first_wgrad_finish_kernel(const double* __restrict__ sums, int C, float* __restrict__ dw /* [C][3] */) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 3 * C) dw[(t % C) * 3 + t / C] = (float)sums[t];
}

unsigned bn_grid(long long rows_out, int rpb, int groups = 1) {
    const long long want = (rows_out + rpb - 1) / rpb;
    const long long cap = ((long long)sm_count() * 8 + groups - 1) / groups;      // CTAs per group
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : (cap > 0 ? cap : 1));
}

int check_shape(const char* who, long long rows, int C, int pool, int groups = 1) {
    POF_REQUIRE(rows > 0 && C >= 4 && (C % 4) == 0 && C <= 1024 && kBnThreads % (C >> 2) == 0, POF_ERR_BAD_SHAPE,
                "%s: need rows > 0 and C in {4..1024} with C / 4 dividing %d (got rows=%lld C=%d)", who, kBnThreads, rows, C);
    POF_REQUIRE(groups >= 1 && groups <= 65535 && rows % groups == 0, POF_ERR_BAD_SHAPE,
                "%s: the rows must split evenly into 1..65535 groups (got rows=%lld groups=%d)", who, rows, groups);
    POF_REQUIRE(pool == 1 || (pool == 2 && (rows / groups) % 2 == 0), POF_ERR_BAD_PARAM,
                "%s: pool must be 1, or 2 with an even number of rows per group", who);
    return POF_OK;
}

}  // namespace
}  // namespace pof

extern "C" {

int pof_bn_act_stats(const float* y, long long rows, int C, int groups, double* sums, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    POF_REQUIRE(y && sums, POF_ERR_NULL_POINTER, "pof_bn_act_stats: null pointer");
    if (int rc = check_shape("pof_bn_act_stats", rows, C, 1, groups)) return rc;
    POF_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, POF_ERR_BAD_PARAM, "pof_bn_act_stats: y must be 16-byte aligned");
    POF_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)groups * C * sizeof(double), stream));
    const int rpb = kBnThreads / (C >> 2);
    const long long rows_g = rows / groups;
    const dim3 grid(bn_grid((rows_g + kBnRowsPerThread - 1) / kBnRowsPerThread, rpb, groups), groups);
    bn_stats_kernel<<<grid, kBnThreads, 2 * (size_t)rpb * C * sizeof(float), stream>>>(y, rows_g, C, sums);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_bn_act_fwd(const float* y, const double* sums, const float* gamma, const float* beta, long long rows, int C, int groups,
                   int pool, float eps, float slope, float momentum, float* z, float* mean, float* invstd, float* running_mean,
                   float* running_var, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    POF_REQUIRE(y && sums && gamma && beta && z && mean && invstd, POF_ERR_NULL_POINTER, "pof_bn_act_fwd: null pointer");
    POF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), POF_ERR_BAD_PARAM, "pof_bn_act_fwd: running_mean and running_var go together");
    if (int rc = check_shape("pof_bn_act_fwd", rows, C, pool, groups)) return rc;
    const uintptr_t al = reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gamma) |
                         reinterpret_cast<uintptr_t>(beta);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_bn_act_fwd: tensors must be 16-byte aligned");
    BnArgs a{};
    a.y = y; a.sums = sums; a.gamma = gamma; a.beta = beta; a.mean = mean; a.invstd = invstd;
    a.running_mean = running_mean; a.running_var = running_var; a.out = z;
    a.rows = rows / groups; a.C = C; a.pool = pool; a.groups = groups; a.eps = eps; a.slope = slope; a.momentum = momentum;
    const dim3 grid(bn_grid(a.rows / pool, kBnThreads / (C >> 2), groups), groups);
    if (pool == 2) bn_act_fwd_kernel<2><<<grid, kBnThreads, 0, stream>>>(a);
    else bn_act_fwd_kernel<1><<<grid, kBnThreads, 0, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_conv_first_wgrad(const float* dy, const float* cutouts, long long M, int P, int C, double* sums, float* dw, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    POF_REQUIRE(dy && cutouts && sums && dw, POF_ERR_NULL_POINTER, "pof_conv_first_wgrad: null pointer");
    POF_REQUIRE(M > 0 && P >= 1, POF_ERR_BAD_SHAPE, "pof_conv_first_wgrad: bad shape M=%lld P=%d", M, P);
    if (int rc = check_shape("pof_conv_first_wgrad", M * P, C, 1)) return rc;
    POF_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0, POF_ERR_BAD_PARAM, "pof_conv_first_wgrad: dy must be 16-byte aligned");
    POF_CUDA(cudaMemsetAsync(sums, 0, 3 * (size_t)C * sizeof(double), stream));
    const long long rows = M * P;
    const int rpb = kBnThreads / (C >> 2);
    const unsigned grid = bn_grid((rows + kBnRowsPerThread - 1) / kBnRowsPerThread, rpb);
    conv_first_wgrad_kernel<<<grid, kBnThreads, 3 * (size_t)rpb * C * sizeof(float), stream>>>(dy, cutouts, rows, P, C, sums);
    POF_CUDA(cudaGetLastError());
    conv_first_wgrad_finish_kernel<<<(3 * C + 127) / 128, 128, 0, stream>>>(sums, C, dw);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_bn_act_bwd(const float* y, const float* dz, const float* mean, const float* invstd, const float* gamma, const float* beta,
                   long long rows, int C, int groups, int pool, float slope, double* sums, float* dx, float* dgamma, float* dbeta,
                   void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    POF_REQUIRE(y && dz && mean && invstd && gamma && beta && sums && dx && dgamma && dbeta, POF_ERR_NULL_POINTER, "pof_bn_act_bwd: null pointer");
    if (int rc = check_shape("pof_bn_act_bwd", rows, C, pool, groups)) return rc;
    const uintptr_t al = reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dx) |
                         reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_bn_act_bwd: tensors must be 16-byte aligned");
    BnArgs a{};
    a.y = y; a.dz = dz; a.sums = sums; a.gamma = gamma; a.beta = beta;
    a.mean = const_cast<float*>(mean); a.invstd = const_cast<float*>(invstd);
    a.out = dx; a.dgamma = dgamma; a.dbeta = dbeta;
    a.rows = rows / groups; a.C = C; a.pool = pool; a.groups = groups; a.slope = slope;
    POF_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)groups * C * sizeof(double), stream));
    const int rpb = kBnThreads / (C >> 2);
    const long long rows_out = a.rows / pool;
    const dim3 grid_r(bn_grid((rows_out + kBnRowsPerThread / pool - 1) / (kBnRowsPerThread / pool), rpb, groups), groups);
    const size_t smem = 2 * (size_t)rpb * C * sizeof(float);
    if (pool == 2) bn_act_bwd_reduce_kernel<2><<<grid_r, kBnThreads, smem, stream>>>(a, sums);
    else bn_act_bwd_reduce_kernel<1><<<grid_r, kBnThreads, smem, stream>>>(a, sums);
    POF_CUDA(cudaGetLastError());
    const dim3 grid(bn_grid(rows_out, rpb, groups), groups);
    if (pool == 2) bn_act_bwd_kernel<2><<<grid, kBnThreads, 0, stream>>>(a);
    else bn_act_bwd_kernel<1><<<grid, kBnThreads, 0, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
