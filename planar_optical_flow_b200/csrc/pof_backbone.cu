// Backbone glue for the streaming engine.
//
// The 1-D convolutions of DROW / SpatialDROW (/root/reference/src/depracted/model/dr_spaam.py:8-12,
// 49-59, 87-114) stay on cuDNN tensor cores.  What PyTorch runs BETWEEN them in eval mode — bias add,
// BatchNorm (folded into the convolution by the engine), LeakyReLU, max-pool — is three to four extra
// passes over multi-GB activations.  These kernels do all of it in ONE pass over channels-last
// activations [rows, C] and, optionally, emit the operand split that lets cuDNN's TF32 tensor cores
// reproduce fp32 accuracy ("3xTF32"):
//
//     x = hi + lo,  hi = round_to_tf32(x),  lo = x - hi           (both exact in fp32)
//     conv(x, w) ~= conv(hi, w_hi) + conv(lo, w_hi) + conv(hi, w_lo)   (dropped lo*lo term: 2^-22)
//
// evaluated as ONE convolution over 3*C input channels [hi | lo | hi] against weights
// [w_hi | w_hi | w_lo]; every product the tensor core forms is exact (11-bit x 11-bit significands)
// and accumulation is fp32.
//
//   pof_act_fwd          y[rows_in, C] (+bias) -> LeakyReLU -> max over `pool` consecutive rows ->
//                        plain [rows_out, C] and/or split [rows_out, 3C]
//   pof_conv_first_fwd   the 1 -> C first layer (k = 3, zero padding) straight from the cutouts, fused
//                        with bias + LeakyReLU + split: cuDNN needs 2.4 ms for these 3 GFLOP.
// Both are pure streaming kernels (HBM bound): 128-bit accesses, grid-stride.
#include <cuda_fp16.h>

#include "pof_common.cuh"

namespace pof {
namespace {

__device__ __forceinline__ float tf32_round(float x) {
    // round to nearest even on the 13 dropped significand bits
    unsigned u = __float_as_uint(x);
    u += 0x0fffu + ((u >> 13) & 1u);
    return __uint_as_float(u & 0xffffe000u);
}
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// parts = 3: [hi | lo | hi] with the exact remainder (operand of a cuDNN TF32 convolution against
//            [w_hi | w_hi | w_lo]);  parts = 2: [hi | lo] with lo rounded to TF32 (operand of pof_conv_tc_fwd).
//            parts = POF_SPLIT_F16: [hi | lo] in binary16 (operand of pof_conv_tc_f16_fwd).
__device__ __forceinline__ void emit(float4 v, size_t row, int c, int C, float* plain, void* split_, int parts, float& amax) {
    if (plain) *reinterpret_cast<float4*>(plain + row * C + c) = v;
    if (split_ && parts == POF_SPLIT_F16) {
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        __half* base = reinterpret_cast<__half*>(split_) + row * 2 * (size_t)C + c;
        *reinterpret_cast<uint2*>(base) = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
        *reinterpret_cast<uint2*>(base + C) = make_uint2(*reinterpret_cast<const unsigned*>(&l01), *reinterpret_cast<const unsigned*>(&l23));
        return;
    }
    float* split = reinterpret_cast<float*>(split_);
    if (split) {
        const float4 hi = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        float* base = split + row * parts * (size_t)C + c;
        st_stream_f4(reinterpret_cast<float4*>(base), hi);
        if (parts == 3) {
            st_stream_f4(reinterpret_cast<float4*>(base + 2 * (size_t)C), hi);
        } else {
            lo = make_float4(tf32_round(lo.x), tf32_round(lo.y), tf32_round(lo.z), tf32_round(lo.w));
        }
        st_stream_f4(reinterpret_cast<float4*>(base + C), lo);
    }
}

template <int POOL>
__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ y, const float* __restrict__ bias, float slope,
                                                  int C, long long rows_out, float* plain, void* split, int parts, int* status) {
    float amax = 0.f;
    const int c4n = C >> 2;
    const long long total = rows_out * c4n;
    const bool small = total < (1ll << 32);             // 32-bit index arithmetic (a 64-bit division costs ~50 issue slots)
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long row = small ? (long long)((unsigned)t / (unsigned)c4n) : t / c4n;
        const int c = (int)(t - row * c4n) << 2;
        const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = ld_stream_f4(reinterpret_cast<const float4*>(y + (size_t)row * POOL * C + c));
        if (POOL == 2) {
            const float4 w = ld_stream_f4(reinterpret_cast<const float4*>(y + ((size_t)row * POOL + 1) * C + c));
            v = make_float4(fmaxf(v.x, w.x), fmaxf(v.y, w.y), fmaxf(v.z, w.z), fmaxf(v.w, w.w));   // lrelu is monotone
        }
        v = make_float4(lrelu(v.x + b.x, slope), lrelu(v.y + b.y, slope), lrelu(v.z + b.z, slope), lrelu(v.w + b.w, slope));
        emit(v, (size_t)row, c, C, plain, split, parts, amax);
    }
    if (status && !(amax <= 65504.f)) atomicCAS(status, 0, 16);      // beyond binary16 (or NaN): the split is not a split
}

// out[m, l, c] = lrelu(b[c] + sum_k w[c, k] * x[m, l + k - 1]),  x = cutouts [M, P], zero padded.
// A thread keeps ONE group of G adjacent channels for all the rows it visits (C/G divides the block size), so its
// weights and biases live in registers; the first version re-read them from shared memory for every row with
// an 8-way bank conflict and ran at 30 % of the write bandwidth.  G = 8 for the float16 split (what the engine
// uses): a thread then stores 16 bytes of hi and 16 bytes of lo per row, and the 8 threads of a 64-channel row write
// two full 128-byte lines; G = 4 (8-byte stores) measured 3.6 TB/s.
template <int G>
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float slope, int P, int C,
                                                         long long rows /* M*P */, float* plain, void* split, int parts, int* status) {
    float amax = 0.f;
    const unsigned cgn = (unsigned)C / G;
    const unsigned rpb = blockDim.x / cgn;             // rows per block and pass
    const int c = (int)(threadIdx.x % cgn) * G;
    float w0[G], w1[G], w2[G], bb[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        w0[j] = __ldg(w + 3 * (c + j)); w1[j] = __ldg(w + 3 * (c + j) + 1); w2[j] = __ldg(w + 3 * (c + j) + 2);
        bb[j] = __ldg(bias + c + j);
    }
    const long long row_first = (long long)blockIdx.x * rpb + threadIdx.x / cgn, row_step = (long long)gridDim.x * rpb;
    int l = (int)(row_first % P);                          // position inside the cutout, advanced without a division per row
    const int l_step = (int)(row_step % P);
    for (long long row = row_first; row < rows; row += row_step, l = l + l_step >= P ? l + l_step - P : l + l_step) {
        const float xc = __ldg(x + row);
        const float xl = l > 0 ? __ldg(x + row - 1) : 0.f;
        const float xr = l < P - 1 ? __ldg(x + row + 1) : 0.f;
        float o[G];
#pragma unroll
        for (int j = 0; j < G; ++j) o[j] = lrelu(fmaf(w2[j], xr, fmaf(w1[j], xc, fmaf(w0[j], xl, bb[j]))), slope);
        if (G == 8 && parts == POF_SPLIT_F16 && split && !plain) {
            unsigned hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 h2 = __floats2half2_rn(o[2 * j], o[2 * j + 1]);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(o[2 * j] - hf.x, o[2 * j + 1] - hf.y);
                amax = fmaxf(amax, fmaxf(fabsf(o[2 * j]), fabsf(o[2 * j + 1])));
                hi[j] = *reinterpret_cast<const unsigned*>(&h2);
                lo[j] = *reinterpret_cast<const unsigned*>(&l2);
            }
            __half* base = reinterpret_cast<__half*>(split) + (size_t)row * 2 * (size_t)C + c;
            st_stream_f4(reinterpret_cast<float4*>(base), make_float4(__uint_as_float(hi[0]), __uint_as_float(hi[1]), __uint_as_float(hi[2]), __uint_as_float(hi[3])));
            st_stream_f4(reinterpret_cast<float4*>(base + C), make_float4(__uint_as_float(lo[0]), __uint_as_float(lo[1]), __uint_as_float(lo[2]), __uint_as_float(lo[3])));
        } else {
#pragma unroll
            for (int j = 0; j < G; j += 4)
                emit(make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]), (size_t)row, c + j, C, plain, split, parts, amax);
        }
    }
    if (status && !(amax <= 65504.f)) atomicCAS(status, 0, 16);
}

// One warp per cutout: y [M, L, C] (raw output of the last convolution, channels last)
//   -> lrelu(y + bias) -> mean over L -> H heads (1x1 convolutions) -> sigmoid on the first n_sig.
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ y, const float* __restrict__ bias, float slope,
                                                   int L, int C, const float* __restrict__ w_head,
                                                   const float* __restrict__ b_head, int H, int n_sig, long long M,
                                                   float* __restrict__ out, float* __restrict__ out_rest) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
        float acc[8];                              // up to 8 heads
#pragma unroll
        for (int h = 0; h < 8; ++h) acc[h] = 0.f;
        for (int c = lane * 4; c < C; c += 128) {
            const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = 0; l < L; ++l) {
                const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(y + ((size_t)m * L + l) * C + c));
                s.x += lrelu(v.x + b.x, slope); s.y += lrelu(v.y + b.y, slope);
                s.z += lrelu(v.z + b.z, slope); s.w += lrelu(v.w + b.w, slope);
            }
            const float fl = (float)L;
            s.x = __fdiv_rn(s.x, fl); s.y = __fdiv_rn(s.y, fl); s.z = __fdiv_rn(s.z, fl); s.w = __fdiv_rn(s.w, fl);
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                if (h < H) {
                    const float4 w = __ldg(reinterpret_cast<const float4*>(w_head + (size_t)h * C + c));
                    acc[h] = fmaf(s.x, w.x, fmaf(s.y, w.y, fmaf(s.z, w.z, fmaf(s.w, w.w, acc[h]))));
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            if (h < H) {
                float v = warp_sum(acc[h]) + __ldg(b_head + h);
                if (h < n_sig) v = 1.f / (1.f + expf(-v));
                if (lane == 0) {
                    if (!out_rest) out[(size_t)m * H + h] = v;
                    else if (h < n_sig) out[(size_t)m * n_sig + h] = v;
                    else out_rest[(size_t)m * (H - n_sig) + (h - n_sig)] = v;
                }
            }
        }
    }
}

unsigned stream_grid(long long items, int threads) {
    const long long want = (items + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace pof

extern "C" {

int pof_act_fwd(const float* y, const float* bias, long long rows_in, int C, int pool, float slope, float* out_plain,
                void* out_split, int split_parts, int* status, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (rows_in == 0) return POF_OK;
    POF_REQUIRE(y && (out_plain || out_split), POF_ERR_NULL_POINTER, "pof_act_fwd: null input or no output");
    POF_REQUIRE(C >= 4 && (C % 4) == 0, POF_ERR_BAD_SHAPE, "pof_act_fwd: C must be a multiple of 4 (got %d)", C);
    POF_REQUIRE(pool == 1 || pool == 2, POF_ERR_UNSUPPORTED, "pof_act_fwd: pool must be 1 or 2 (got %d)", pool);
    POF_REQUIRE(split_parts == 2 || split_parts == 3 || split_parts == POF_SPLIT_F16, POF_ERR_BAD_PARAM,
                "pof_act_fwd: split_parts must be 2, 3 or POF_SPLIT_F16");
    POF_REQUIRE(rows_in > 0 && rows_in % pool == 0, POF_ERR_BAD_SHAPE, "pof_act_fwd: rows_in must be a positive multiple of pool");
    const uintptr_t al = reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) |
                         reinterpret_cast<uintptr_t>(out_plain) | reinterpret_cast<uintptr_t>(out_split);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_act_fwd: tensors must be 16-byte aligned");
    const long long rows_out = rows_in / pool;
    const unsigned grid = stream_grid(rows_out * (C >> 2), 256);
    if (pool == 1) act_kernel<1><<<grid, 256, 0, stream>>>(y, bias, slope, C, rows_out, out_plain, out_split, split_parts, status);
    else act_kernel<2><<<grid, 256, 0, stream>>>(y, bias, slope, C, rows_out, out_plain, out_split, split_parts, status);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_conv_first_fwd(const float* cutouts, const float* weight, const float* bias, long long M, int P, int C, float slope,
                       float* out_plain, void* out_split, int split_parts, int* status, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (M == 0) return POF_OK;
    POF_REQUIRE(cutouts && weight && bias && (out_plain || out_split), POF_ERR_NULL_POINTER, "pof_conv_first_fwd: null pointer");
    POF_REQUIRE(M > 0 && P >= 1 && C >= 4 && (C % 4) == 0 && C <= 1024, POF_ERR_BAD_SHAPE,
                "pof_conv_first_fwd: bad shape M=%lld P=%d C=%d", M, P, C);
    const uintptr_t al = reinterpret_cast<uintptr_t>(out_plain) | reinterpret_cast<uintptr_t>(out_split);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_conv_first_fwd: outputs must be 16-byte aligned");
    POF_REQUIRE(split_parts == 2 || split_parts == 3 || split_parts == POF_SPLIT_F16, POF_ERR_BAD_PARAM,
                "pof_conv_first_fwd: split_parts must be 2, 3 or POF_SPLIT_F16");
    const long long rows = M * P;
    POF_REQUIRE(256 % (C >> 2) == 0, POF_ERR_BAD_SHAPE, "pof_conv_first_fwd: C / 4 must divide 256 (got C = %d)", C);
    if (C % 8 == 0 && 256 % (C >> 3) == 0 && split_parts == POF_SPLIT_F16 && out_split && !out_plain) {
        const unsigned grid = stream_grid(rows * (C >> 3), 256);
        conv_first_kernel<8><<<grid, 256, 0, stream>>>(cutouts, weight, bias, slope, P, C, rows, out_plain, out_split, split_parts, status);
    } else {
        const unsigned grid = stream_grid(rows * (C >> 2), 256);
        conv_first_kernel<4><<<grid, 256, 0, stream>>>(cutouts, weight, bias, slope, P, C, rows, out_plain, out_split, split_parts, status);
    }
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_head_fwd(const float* y, const float* bias, long long M, int L, int C, float slope, const float* w_head,
                 const float* b_head, int H, int n_sigmoid, float* out, float* out_rest, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (M == 0) return POF_OK;
    POF_REQUIRE(y && w_head && b_head && out, POF_ERR_NULL_POINTER, "pof_head_fwd: null pointer");
    POF_REQUIRE(M > 0 && L >= 1 && C >= 4 && (C % 4) == 0, POF_ERR_BAD_SHAPE, "pof_head_fwd: bad shape M=%lld L=%d C=%d", M, L, C);
    POF_REQUIRE(H >= 1 && H <= 8 && n_sigmoid >= 0 && n_sigmoid <= H, POF_ERR_BAD_SHAPE,
                "pof_head_fwd: 1..8 heads supported (got %d, %d with sigmoid)", H, n_sigmoid);
    const uintptr_t al = reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(w_head);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_head_fwd: y, bias and w_head must be 16-byte aligned");
    const unsigned grid = stream_grid(M * 32, 256);
    head_kernel<<<grid, 256, 0, stream>>>(y, bias, slope, L, C, w_head, b_head, H, n_sigmoid, M, out, out_rest);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
