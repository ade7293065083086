// Kernel 2 — auto-regressive spatial-attention memory update (forward).
//
// Replaces everything after the two embedding convolutions in
// _SpatialAttention.forward (/root/reference/src/depracted/model/dr_spaam.py:183-215):
// the reference forms a dense [N, N] similarity, masks it to the +-hw window,
// soft-maxes and multiplies the dense weights with the [N, C*L] template
// (8.5 GFLOP of multiplications by zero per JRDB scan).  Here one fused kernel
// does the windowed form:
//
//   sim[i,k]   = <emb_x[i], emb_t[clamp(i-hw+k)]>             k = 0..W-1     (:184,:187)
//   w[i,.]     = softmax over the UNIQUE in-range neighbours                  (:197-201)
//   out[i,:]   = alpha*x[i,:] + (1-alpha) * sum_k w[i,k] * tmpl[i-hw+k,:]     (:210-215)
//
// HBM-bound: per point it must read x (14 KB) and the template row (14 KB) and
// write out (14 KB).  Layout of the work on B200:
//   * a CTA owns (sequence b, a chunk of consecutive points, a slice of 4*T
//     channels, T = 896 = a whole DR-SPAAM row); thread t owns 4 adjacent channels (one float4) and MARCHES
//     along the chunk's points;
//   * the W template rows a point needs are kept in a per-thread circular
//     REGISTER window (W float4), so each template element is fetched once per
//     chunk (plus the 2*hw halo rows at the chunk ends, which neighbouring CTAs
//     hit in L2) and the 11-fold neighbour re-use costs no memory traffic at all;
//     the loop is unrolled by the window size so every window slot is a fixed
//     register;
//   * rows arrive through a shared-memory RING filled by TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP): one elected thread
//     keeps kGateRing stages (template row + x row) in flight per CTA.  A first
//     version prefetched with plain LDG into registers and stalled at 51 % of HBM
//     peak on long-scoreboard waits (loads share 6 scoreboard slots per warp, so a
//     deeper register prefetch does not add memory-level parallelism); the ring
//     decouples bytes in flight from registers and occupancy;
//   * the similarities and soft-max weights of the chunk are computed once in a
//     prologue (one warp per point: neighbour embeddings are 512-byte rows read
//     as one float4 per lane, dot products finished with warp shuffles) while the
//     first ring stages are already in flight, and are broadcast from shared
//     memory in the streaming loop.
#include <cuda_fp16.h>
#include <math.h>

#include "pof_common.cuh"

namespace pof {
namespace {

constexpr int kMaxChunk = 160;     // points per CTA (upper bound, sizes the weight table)
#ifndef POF_GATE_THREADS
#define POF_GATE_THREADS 896
#endif
#ifndef POF_GATE_MINBLOCKS
#define POF_GATE_MINBLOCKS 1
#endif
#ifndef POF_GATE_RING
#define POF_GATE_RING 6
#endif
#ifndef POF_GATE_CHUNK_TARGET
#define POF_GATE_CHUNK_TARGET 128
#endif
// 896 threads x float4 = 3584 channels per CTA: one CTA streams WHOLE DR-SPAAM rows, so the
// similarity prologue runs once per point (narrower slices repeat it per slice); one CTA per SM
// with kGateRing = 6 stages (28 KB each: template row + x row) = 168 KB of HBM reads in flight
// per SM.  Measured on B200 (tools/tune_gate.py, 128 JRDB sequences, % of the measured 6547 GB/s
// copy peak): LDG-prefetch version 51 %; TMA ring T128/R8 67 %, T224/R8 71 %, T256/R8 79 %,
// T448x2/R6 83 %, T448/R12 89 %, T896/R6 96 %.
constexpr int kGateThreads = POF_GATE_THREADS;
constexpr int kGateMinBlocks = POF_GATE_MINBLOCKS;
constexpr int kGateRing = POF_GATE_RING;   // stages (template row + x row) kept in flight per CTA by TMA

struct GateArgs {
    const float* x;
    const float* tmpl;
    const float* emb_x;
    const float* emb_t;
    float* out;
    float* out2;         // backward only: g_x
    float* feat_fused;
    float* attn_w;
    int B, N, CL, E;
    int chunk_len, n_chunks;
    float alpha, beta;   // beta = (float)(1.0 - alpha)
    const float* g_feat; // MODE 2: gradient of feat_fused [B, N, W] or null
    float* g_s;          // MODE 2: gradient of the similarities [B, N, W]
    void* out_split;     // forward only, optional: the new memory as binary16 [hi | lo] rows of split_c channels
    int split_c;         //   (operand of pof_conv_tc_f16_fwd: row r = point * (CL / split_c) + l is [hi(split_c) | lo(split_c)])
    int* status;         // optional device int: 32 = a ring wait timed out, 16 = a memory value left the binary16 range
};

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

// Similarities + soft-max weights for the chunk's points; one warp per point.
template <int W>
__device__ __forceinline__ void chunk_weights(const GateArgs& a, int b, int i0, int len, bool write_global,
                                              float (*w_s)[(W + 3) & ~3]) {
    constexpr int HW = W / 2;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    const float* ex_b = a.emb_x + (size_t)b * a.N * a.E;
    const float* et_b = a.emb_t + (size_t)b * a.N * a.E;
    for (int p = warp; p < len; p += n_warps) {
        const int i = i0 + p;
        float part[W];
#pragma unroll
        for (int k = 0; k < W; ++k) part[k] = 0.f;
        for (int e = lane * 4; e < a.E; e += 128) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(ex_b + (size_t)i * a.E + e));
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int j = min(max(i - HW + k, 0), a.N - 1);      // clamped neighbour (:152)
                const float4 t = __ldg(reinterpret_cast<const float4*>(et_b + (size_t)j * a.E + e));
                part[k] = fmaf(q.x, t.x, fmaf(q.y, t.y, fmaf(q.z, t.z, fmaf(q.w, t.w, part[k]))));
            }
        }
        float mine = 0.f;   // lane k ends up owning similarity k
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const float s = warp_sum(part[k]);
            if (lane == k) mine = s;
        }
        const int j_raw = i - HW + lane;
        const bool in_window = lane < W && j_raw >= 0 && j_raw <= a.N - 1;
        const float top = warp_max(in_window ? mine : -INFINITY);
        const float ex = in_window ? expf(mine - top) : 0.f;
        const float wgt = ex / warp_sum(ex);
        if (lane < W) {
            w_s[p][lane] = wgt;
            if (write_global) {
                const size_t o = ((size_t)b * a.N + i) * W + lane;
                a.feat_fused[o] = mine;
                if (a.attn_w) a.attn_w[o] = wgt;
            }
        }
    }
}

// Transposed weights for the backward stream: g_tmpl[j] = beta * sum_k' wT[j][k'] * g_out[j-hw+k']
// with wT[j][k'] = w[j-hw+k'][W-1-k'] (0 when that point does not exist).
template <int W>
__device__ __forceinline__ void chunk_weights_transposed(const GateArgs& a, int b, int i0, int len,
                                                         float (*w_s)[(W + 3) & ~3]) {
    constexpr int HW = W / 2;
    const float* wb = a.attn_w + (size_t)b * a.N * W;
    for (int t = threadIdx.x; t < len * W; t += blockDim.x) {
        const int p = t / W, k = t - p * W;
        const int i = i0 + p - HW + k;
        w_s[p][k] = (i >= 0 && i < a.N) ? __ldg(wb + (size_t)i * W + (W - 1 - k)) : 0.f;
    }
}

// ---- async-proxy helpers (TMA bulk copies completing on an mbarrier) --------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded like every wait of the convolution kernel: a copy that never lands (a bad pointer) becomes status 32 and
// garbage in the output, never a hung GPU.  The fast path is one try_wait (which itself blocks for a hardware time slice).
constexpr long long kGateWaitLimit = 2000000000LL;     // ~1 s of SM clocks
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity, int* status) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 64; ++i)
            if (mbar_try(bar, parity)) return;
        if (clock64() - t0 > kGateWaitLimit) {
            if (status) atomicCAS(status, 0, 32);
            return;
        }
    }
}
// 1-D bulk copy global -> shared (SASS: UBLKCP), bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// MODE 0 (forward):  out[i]  = alpha*x[i] + beta * sum_k w[i][k]  * tmpl[i-hw+k]
// MODE 1 (backward): out[j]  = beta * sum_k wT[j][k] * g_out[j-hw+k]   (g_tmpl; `tmpl` = g_out)
//                    out2[j] = alpha * g_out[j]                         (g_x)
// MODE 2 (backward): g_w[i][k] = beta * <g_out[i], tmpl[i-hw+k]>  (`x` = g_out), then the soft-max Jacobian and the direct
//                    feat_fused gradient: g_s[i][k] = w[i][k] * (g_w[i][k] - sum_k' w[i][k'] g_w[i][k']) + g_feat[i][k].
//                    The template rows come through the same ring and register window as in the forward, so each is
//                    fetched once per chunk (the first version of this pass - one warp per point, gate_bwd_scores_kernel -
//                    re-read the W rows of every point out of L2).  A thread's four channels give W partial products per
//                    point; a warp folds its 32 x W partials with a transposing butterfly (16 shuffles for up to 16 values
//                    instead of 5 per value), the 28 warp sums of a point meet in shared memory, and warp 0 finishes the point
//                    one stage later, behind the barrier the ring needs anyway.  Clamped duplicates at the sequence ends
//                    have weight 0 in w, so their products never reach g_s.
//
// The chunk is consumed as a sequence of STAGES s = 0 .. len+W-2.  Stage s carries template row
// clamp(i0-hw+s) and, once the window is full (s >= W-1), the x row of point p = s-(W-1).  One
// elected thread keeps kGateRing stages in flight with TMA bulk copies into a shared-memory ring
// (completion on one mbarrier per slot); every thread then moves ITS 16 bytes of the new template
// row into its register window, so shared memory is only a deep prefetch queue (1 write + 1 read
// per element) and the W-fold neighbour re-use still costs no memory traffic.
// floats of the per-chunk weight table; MODE 2 keeps two points' per-warp partial sums there instead
template <int W>
struct GateTable {
    static constexpr size_t kWeights = (size_t)kMaxChunk * ((W + 3) & ~3), kPartials = (size_t)2 * (kGateThreads / 32) * 16;
    static constexpr size_t kFloats = kWeights > kPartials ? kWeights : kPartials;
};

// Sum 16 per-lane values over the 32 lanes of a warp with a transposing butterfly: after the five steps lane L holds the
// warp total of value index  ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)  (both lanes of a pair hold it).
__device__ __forceinline__ float warp_fold16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = (lane & 16) != 0;
        const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = (lane & 8) != 0;
        const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = (lane & 4) != 0;
        const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const bool up = (lane & 2) != 0;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int W, int MODE>
__global__ void __launch_bounds__(kGateThreads, kGateMinBlocks) gate_stream_kernel(const GateArgs a) {
    constexpr int HW = W / 2;
    constexpr int WPAD = (W + 3) & ~3;
    constexpr int R = kGateRing;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [R][2][T] float4 stage buffers | weights [kMaxChunk][WPAD] | R mbarriers
    float4* stage = reinterpret_cast<float4*>(smem_raw);
    float(*w_s)[WPAD] = reinterpret_cast<float(*)[WPAD]>(smem_raw + (size_t)R * 2 * kGateThreads * sizeof(float4));
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)R * 2 * kGateThreads * sizeof(float4) +
                                                                     sizeof(float) * GateTable<W>::kFloats);

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int i0 = blockIdx.x * a.chunk_len;
    const int len = min(a.chunk_len, a.N - i0);
    const int n_stages = len + W - 1;
    const int ch0 = blockIdx.y * kGateThreads * 4;                    // first channel of this slice
    const int ch = ch0 + tid * 4;
    const bool active = ch < a.CL;
    const unsigned row_bytes = (unsigned)(min(kGateThreads * 4, a.CL - ch0) * sizeof(float));
    const size_t seq = (size_t)b * a.N * a.CL;
    const size_t CL = (size_t)a.CL;
    const float* t_base = a.tmpl + seq + ch0;
    const float* x_base = a.x + seq + ch0;
    float* o_col = a.out + seq + (active ? ch : 0);

    auto issue = [&](int s) {        // thread 0 only
        const int slot = s % R;
        const bool with_x = (MODE == 0 || MODE == 2) && s >= W - 1;
        mbar_expect_tx(&full[slot], with_x ? 2 * row_bytes : row_bytes);
        const int r = min(max(i0 - HW + s, 0), a.N - 1);              // clamped like the reference's table (:152)
        bulk_g2s(stage + (size_t)slot * 2 * kGateThreads, t_base + r * CL, row_bytes, &full[slot]);
        if (with_x)
            bulk_g2s(stage + ((size_t)slot * 2 + 1) * kGateThreads, x_base + (size_t)(i0 + s - (W - 1)) * CL, row_bytes,
                     &full[slot]);
    };

    if (tid == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) mbar_init(&full[r], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < min(R, n_stages); ++s) issue(s);          // in flight while the weights are computed
    }
    if (MODE == 0) chunk_weights<W>(a, b, i0, len, blockIdx.y == 0, w_s);
    else if (MODE == 1) chunk_weights_transposed<W>(a, b, i0, len, w_s);
    __syncthreads();
    // MODE 2: the weight table's memory holds the per-warp partial sums of two points in flight: red[parity][warp][16]
    float* red = reinterpret_cast<float*>(w_s);
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kGateThreads / 32;
    static_assert(W <= 16 && 2 * kWarps * 16 <= GateTable<W>::kFloats, "partial sums do not fit the weight table");
    // MODE 2: warp 0 turns the warp sums of point p into g_s[p]
    auto finish_point = [&](int p) {
        const int i = i0 + p;
        float gw = 0.f;
        if (lane < W) {
            const float* r = red + (size_t)(p & 1) * kWarps * 16 + lane;
#pragma unroll 4
            for (int w = 0; w < kWarps; ++w) gw += r[w * 16];
            gw *= a.beta;
        }
        const size_t o = ((size_t)b * a.N + i) * W + lane;
        const float wv = lane < W ? __ldg(a.attn_w + o) : 0.f;
        const float mean = warp_sum(wv * gw);
        if (lane < W) a.g_s[o] = wv * (gw - mean) + (a.g_feat ? __ldg(a.g_feat + o) : 0.f);
    };

    float4 win[W];
#pragma unroll
    for (int k = 0; k < W; ++k) win[k] = f4_zero();
    const float alpha = a.alpha, beta = a.beta;
    // optional second output (forward): the same row as the binary16 [hi | lo] operand of the convolutions that consume
    // the new memory (the gate embedding and conv block 3), written in the pass that already holds it in registers
    __half* s_col = nullptr;
    float amax = 0.f;
    if (MODE == 0 && a.out_split && active) {
        const int l = ch / a.split_c, c = ch - l * a.split_c;
        s_col = reinterpret_cast<__half*>(a.out_split) + (size_t)b * a.N * 2 * CL + (size_t)l * 2 * a.split_c + c;
    }

    for (int s0 = 0; s0 < n_stages; s0 += W) {
#pragma unroll
        for (int u = 0; u < W; ++u) {
            const int s = s0 + u;
            if (s < n_stages) {
                const int slot = s % R;
                mbar_wait(&full[slot], (unsigned)(s / R) & 1u, a.status);
                const float4 tv = stage[(size_t)slot * 2 * kGateThreads + tid];
                float4 xv = f4_zero();
                if ((MODE == 0 || MODE == 2) && s >= W - 1) xv = stage[((size_t)slot * 2 + 1) * kGateThreads + tid];
                __syncthreads();                                       // slot fully read: refill it
                if (tid == 0 && s + R < n_stages) issue(s + R);
                win[u] = tv;                                           // row s replaces row s-W
                if (MODE == 2) {
                    // the barrier above also published the warp sums of the previous point
                    if (warp == 0 && s >= W) finish_point(s - W);
                    if (s >= W - 1) {
                        const int p = s - (W - 1);
                        float part[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            if (k < W) {
                                const float4 t = win[(u + 1 + k) % W];                       // template row i-hw+k
                                part[k] = active ? fmaf(xv.x, t.x, fmaf(xv.y, t.y, fmaf(xv.z, t.z, xv.w * t.w))) : 0.f;
                            } else {
                                part[k] = 0.f;
                            }
                        }
                        const float tot = warp_fold16(part, lane);
                        const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        if ((lane & 1) == 0) red[((size_t)(p & 1) * kWarps + warp) * 16 + idx] = tot;
                    }
                } else if (s >= W - 1) {
                    const int p = s - (W - 1);
                    const int i = i0 + p;
                    float wk[WPAD];
#pragma unroll
                    for (int k4 = 0; k4 < WPAD; k4 += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&w_s[p][k4]);
                        wk[k4] = w4.x; wk[k4 + 1] = w4.y; wk[k4 + 2] = w4.z; wk[k4 + 3] = w4.w;
                    }
                    float4 acc = f4_zero();
#pragma unroll
                    for (int k = 0; k < W; ++k) f4_fma(acc, wk[k], win[(u + 1 + k) % W]);   // row p+k
                    if (MODE == 0) {
                        float4 o;
                        o.x = fmaf(alpha, xv.x, beta * acc.x);
                        o.y = fmaf(alpha, xv.y, beta * acc.y);
                        o.z = fmaf(alpha, xv.z, beta * acc.z);
                        o.w = fmaf(alpha, xv.w, beta * acc.w);
                        if (active) st_stream_f4(reinterpret_cast<float4*>(o_col + (size_t)i * CL), o);
                        if (s_col) {
                            const __half2 h01 = __floats2half2_rn(o.x, o.y), h23 = __floats2half2_rn(o.z, o.w);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn(o.x - f01.x, o.y - f01.y), l23 = __floats2half2_rn(o.z - f23.x, o.w - f23.y);
                            __half* dst = s_col + (size_t)i * 2 * CL;
                            *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
                            *reinterpret_cast<uint2*>(dst + a.split_c) = make_uint2(*reinterpret_cast<const unsigned*>(&l01), *reinterpret_cast<const unsigned*>(&l23));
                            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
                        }
                    } else {
                        const float4 c = win[(u + 1 + HW) % W];          // g_out[j] itself
                        if (active) {
                            st_stream_f4(reinterpret_cast<float4*>(o_col + (size_t)i * CL),
                                         make_float4(beta * acc.x, beta * acc.y, beta * acc.z, beta * acc.w));
                            st_stream_f4(reinterpret_cast<float4*>(a.out2 + seq + ch + (size_t)i * CL),
                                         make_float4(alpha * c.x, alpha * c.y, alpha * c.z, alpha * c.w));
                        }
                    }
                }
            }
        }
    }
    if (MODE == 0 && s_col && !(amax <= 65504.f) && a.status) atomicCAS(a.status, 0, 16);   // beyond binary16 (or NaN)
    if (MODE == 2) {                                                   // the last point of the chunk
        __syncthreads();
        if (warp == 0) finish_point(len - 1);
    }
}

template <int W>
constexpr size_t gate_smem_bytes() {
    return (size_t)kGateRing * 2 * kGateThreads * sizeof(float4) + sizeof(float) * GateTable<W>::kFloats +
           kGateRing * sizeof(unsigned long long);
}

template <int W, int MODE>
int launch_gate_stream(const GateArgs& a, dim3 grid, int threads, cudaStream_t stream) {
    constexpr size_t smem = gate_smem_bytes<W>();
    static bool attr_set[64] = {false};          // once per device (and never inside a stream capture after the first call)
    int dev = 0;
    POF_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        POF_CUDA(cudaFuncSetAttribute(gate_stream_kernel<W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) attr_set[dev] = true;
    }
    gate_stream_kernel<W, MODE><<<grid, threads, smem, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

template <int MODE>
int dispatch_gate_stream(int W, const GateArgs& a, dim3 grid, int threads, cudaStream_t stream) {
    switch (W) {
        case 1: return launch_gate_stream<1, MODE>(a, grid, threads, stream);
        case 3: return launch_gate_stream<3, MODE>(a, grid, threads, stream);
        case 5: return launch_gate_stream<5, MODE>(a, grid, threads, stream);
        case 7: return launch_gate_stream<7, MODE>(a, grid, threads, stream);
        case 9: return launch_gate_stream<9, MODE>(a, grid, threads, stream);
        case 11: return launch_gate_stream<11, MODE>(a, grid, threads, stream);
        case 13: return launch_gate_stream<13, MODE>(a, grid, threads, stream);
        case 15: return launch_gate_stream<15, MODE>(a, grid, threads, stream);
    }
    set_error("spaam gate: unsupported window %d", W);
    return POF_ERR_UNSUPPORTED;
}

// Chunking of the points of one sequence: <= kMaxChunk points per CTA, and short
// enough that the grid covers the machine (3 CTAs/SM) about 4 times when the
// batch is small, but never below 16 points (halo overhead 2*hw/len).
void plan_chunks(int B, int N, int CL, GateArgs& a, dim3& grid) {
    const int slices = (CL + kGateThreads * 4 - 1) / (kGateThreads * 4);
    const long long want_ctas = 4ll * kGateMinBlocks * sm_count();
    int n_chunks = (N + POF_GATE_CHUNK_TARGET - 1) / POF_GATE_CHUNK_TARGET;
    const long long per_chunk = (long long)B * slices;
    if (per_chunk * n_chunks < want_ctas) n_chunks = (int)((want_ctas + per_chunk - 1) / per_chunk);
    n_chunks = max(1, min(n_chunks, (N + 15) / 16));
    int chunk_len = (N + n_chunks - 1) / n_chunks;
    if (chunk_len > kMaxChunk) chunk_len = kMaxChunk;
    n_chunks = (N + chunk_len - 1) / chunk_len;
    a.chunk_len = chunk_len;
    a.n_chunks = n_chunks;
    grid = dim3((unsigned)n_chunks, (unsigned)slices, (unsigned)B);
}

// ---- backward helpers -------------------------------------------------------------------------
// g_w[i,k] = beta * <g_out[i], tmpl[i-hw+k]>, then the soft-max Jacobian and the direct
// feat_fused gradient:  g_s[i,k] = w[i,k] * (g_w[i,k] - sum_k' w[i,k'] g_w[i,k']) + g_feat[i,k].
// One warp per point; template rows are re-read by neighbouring warps out of L2.
template <int W>
__global__ void __launch_bounds__(256) gate_bwd_scores_kernel(const GateArgs a, const float* __restrict__ g_out,
                                                              const float* __restrict__ g_feat, float* __restrict__ g_s) {
    constexpr int HW = W / 2;
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= (long long)a.B * a.N) return;
    const int b = (int)(pt / a.N), i = (int)(pt - (long long)b * a.N);
    const size_t seq = (size_t)b * a.N * a.CL;
    const float4* g4 = reinterpret_cast<const float4*>(g_out + seq + (size_t)i * a.CL);
    float part[W];
#pragma unroll
    for (int k = 0; k < W; ++k) part[k] = 0.f;
    for (int c = lane; c < a.CL / 4; c += 32) {
        const float4 g = __ldg(g4 + c);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int j = i - HW + k;
            if (j >= 0 && j < a.N) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(a.tmpl + seq + (size_t)j * a.CL) + c);
                part[k] = fmaf(g.x, t.x, fmaf(g.y, t.y, fmaf(g.z, t.z, fmaf(g.w, t.w, part[k]))));
            }
        }
    }
    float gw = 0.f;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        const float s = warp_sum(part[k]);
        if (lane == k) gw = a.beta * s;
    }
    const size_t o = ((size_t)b * a.N + i) * W + lane;
    const float w = lane < W ? a.attn_w[o] : 0.f;
    const float mean = warp_sum(w * gw);
    if (lane < W) g_s[o] = w * (gw - mean) + (g_feat ? g_feat[o] : 0.f);
}

// g_emb_x[i] = sum_k g_s[i,k] emb_t[clamp(i-hw+k)];  g_emb_t[j] = sum_{(i,k): clamp(i-hw+k)=j} g_s[i,k] emb_x[i].
template <int W>
__global__ void __launch_bounds__(256) gate_bwd_embed_kernel(const GateArgs a, const float* __restrict__ g_s,
                                                             float* __restrict__ g_ex, float* __restrict__ g_et) {
    constexpr int HW = W / 2;
    const int lane = threadIdx.x & 31;
    const long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pt >= (long long)a.B * a.N) return;
    const int b = (int)(pt / a.N), i = (int)(pt - (long long)b * a.N);
    const int N = a.N, E = a.E;
    const float* ex = a.emb_x + (size_t)b * N * E;
    const float* et = a.emb_t + (size_t)b * N * E;
    const float* gs = g_s + (size_t)b * N * W;
    for (int e = lane * 4; e < E; e += 128) {
        float4 acc = f4_zero();
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int j = min(max(i - HW + k, 0), N - 1);
            f4_fma(acc, gs[(size_t)i * W + k], __ldg(reinterpret_cast<const float4*>(et + (size_t)j * E + e)));
        }
        *reinterpret_cast<float4*>(g_ex + ((size_t)b * N + i) * E + e) = acc;

        // transposed: this point as template row j = i
        const int j = i;
        float4 acc_t = f4_zero();
        for (int q = max(0, j - HW); q <= min(N - 1, j + HW); ++q)
            f4_fma(acc_t, gs[(size_t)q * W + (j - q + HW)], __ldg(reinterpret_cast<const float4*>(ex + (size_t)q * E + e)));
        if (j == 0)          // clamped-low duplicates: q - hw + k < 0
            for (int q = 0; q < min(HW, N); ++q)
                for (int k = 0; k < HW - q; ++k)
                    f4_fma(acc_t, gs[(size_t)q * W + k], __ldg(reinterpret_cast<const float4*>(ex + (size_t)q * E + e)));
        if (j == N - 1)      // clamped-high duplicates: q - hw + k > N-1
            for (int q = max(0, N - HW); q < N; ++q)
                for (int k = N - q + HW; k < W; ++k)
                    f4_fma(acc_t, gs[(size_t)q * W + k], __ldg(reinterpret_cast<const float4*>(ex + (size_t)q * E + e)));
        *reinterpret_cast<float4*>(g_et + ((size_t)b * N + i) * E + e) = acc_t;
    }
}

template <int W>
int launch_gate_bwd_small(const GateArgs& a, const float* g_out, const float* g_feat, float* g_s, float* g_ex, float* g_et,
                          cudaStream_t stream) {
    const long long pts = (long long)a.B * a.N;
    const unsigned grid = (unsigned)((pts + 7) / 8);
#ifdef POF_GATE_BWD_SCORES_PER_POINT          // the first version: one warp per point, template rows re-read from L2
    gate_bwd_scores_kernel<W><<<grid, 256, 0, stream>>>(a, g_out, g_feat, g_s);
    POF_CUDA(cudaGetLastError());
#else
    {
        GateArgs sc = a;
        sc.x = g_out;              // the "x row" of a stage carries g_out[i]
        sc.g_feat = g_feat;
        sc.g_s = g_s;
        dim3 sgrid;
        plan_chunks(a.B, a.N, a.CL, sc, sgrid);
        if (sgrid.y != 1) {        // rows wider than one CTA's 3584 channels: the per-point kernel sums over the whole row
            gate_bwd_scores_kernel<W><<<grid, 256, 0, stream>>>(a, g_out, g_feat, g_s);
            POF_CUDA(cudaGetLastError());
        } else if (int rc = launch_gate_stream<W, 2>(sc, sgrid, kGateThreads, stream)) {
            return rc;
        }
    }
#endif
    gate_bwd_embed_kernel<W><<<grid, 256, 0, stream>>>(a, g_s, g_ex, g_et);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // namespace
}  // namespace pof

extern "C" {

int pof_spaam_gate_fwd(const float* x, const float* tmpl, const float* emb_x, const float* emb_t, int B, int N, int CL,
                       int E, int W, float alpha, float* out_tmpl, float* feat_fused, float* attn_w, void* out_split,
                       int split_channels, int* status, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(x && tmpl && emb_x && emb_t && out_tmpl && feat_fused, POF_ERR_NULL_POINTER,
                "pof_spaam_gate_fwd: null tensor pointer");
    POF_REQUIRE(B >= 0 && N >= 1 && CL >= 4 && E >= 4, POF_ERR_BAD_SHAPE, "pof_spaam_gate_fwd: bad shape B=%d N=%d CL=%d E=%d", B,
                N, CL, E);
    POF_REQUIRE((CL % 4) == 0 && (E % 4) == 0, POF_ERR_BAD_SHAPE, "pof_spaam_gate_fwd: CL and E must be multiples of 4");
    POF_REQUIRE(W >= 1 && (W & 1) && W <= 15, POF_ERR_UNSUPPORTED,
                "pof_spaam_gate_fwd: window must be odd and <= 15 (got %d)", W);
    POF_REQUIRE(out_tmpl != tmpl && out_tmpl != x, POF_ERR_BAD_PARAM,
                "pof_spaam_gate_fwd: out_tmpl must not alias x or tmpl (neighbouring points re-read template rows)");
    const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(tmpl) |
                         reinterpret_cast<uintptr_t>(emb_x) | reinterpret_cast<uintptr_t>(emb_t) |
                         reinterpret_cast<uintptr_t>(out_tmpl);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_spaam_gate_fwd: tensors must be 16-byte aligned");
    POF_REQUIRE(B <= 65535, POF_ERR_BAD_SHAPE, "pof_spaam_gate_fwd: B must be <= 65535");
    if (out_split) {
        POF_REQUIRE(split_channels >= 4 && split_channels % 4 == 0 && CL % split_channels == 0, POF_ERR_BAD_SHAPE,
                    "pof_spaam_gate_fwd: split_channels must be a multiple of 4 that divides CL (got %d for CL = %d)", split_channels, CL);
        POF_REQUIRE((reinterpret_cast<uintptr_t>(out_split) & 15) == 0, POF_ERR_BAD_PARAM, "pof_spaam_gate_fwd: out_split must be 16-byte aligned");
    }

    GateArgs a;
    a.x = x; a.tmpl = tmpl; a.emb_x = emb_x; a.emb_t = emb_t;
    a.out = out_tmpl; a.feat_fused = feat_fused; a.attn_w = attn_w;
    a.out_split = out_split; a.split_c = out_split ? split_channels : 0; a.status = status;
    a.g_feat = nullptr; a.g_s = nullptr;
    a.B = B; a.N = N; a.CL = CL; a.E = E;
    a.alpha = alpha;
    a.beta = (float)(1.0 - (double)alpha);

    a.out2 = nullptr;
    dim3 grid;
    plan_chunks(B, N, CL, a, grid);
    return dispatch_gate_stream<0>(W, a, grid, kGateThreads, stream);
}

size_t pof_spaam_gate_bwd_ws_bytes(int B, int N, int W) {
    return (B > 0 && N > 0 && W > 0) ? (size_t)B * N * W * sizeof(float) : 0;
}

int pof_spaam_gate_bwd(const float* tmpl, const float* emb_x, const float* emb_t, const float* attn_w, const float* g_out,
                       const float* g_feat, int B, int N, int CL, int E, int W, float alpha, float* g_x, float* g_tmpl,
                       float* g_emb_x, float* g_emb_t, void* ws, size_t ws_bytes, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(tmpl && emb_x && emb_t && attn_w && g_out && g_x && g_tmpl && g_emb_x && g_emb_t, POF_ERR_NULL_POINTER,
                "pof_spaam_gate_bwd: null tensor pointer");
    POF_REQUIRE(B >= 0 && N >= 1 && CL >= 4 && E >= 4 && (CL % 4) == 0 && (E % 4) == 0, POF_ERR_BAD_SHAPE,
                "pof_spaam_gate_bwd: bad shape B=%d N=%d CL=%d E=%d", B, N, CL, E);
    POF_REQUIRE(W >= 1 && (W & 1) && W <= 15, POF_ERR_UNSUPPORTED, "pof_spaam_gate_bwd: window must be odd and <= 15 (got %d)", W);
    POF_REQUIRE(B <= 65535, POF_ERR_BAD_SHAPE, "pof_spaam_gate_bwd: B must be <= 65535");
    POF_REQUIRE(g_tmpl != g_out && g_x != g_out, POF_ERR_BAD_PARAM, "pof_spaam_gate_bwd: gradients must not alias g_out");
    if (B == 0) return POF_OK;
    POF_REQUIRE(ws && ws_bytes >= pof_spaam_gate_bwd_ws_bytes(B, N, W), POF_ERR_WORKSPACE,
                "pof_spaam_gate_bwd: workspace too small (%zu < %zu)", ws_bytes, pof_spaam_gate_bwd_ws_bytes(B, N, W));

    GateArgs a;
    a.x = nullptr; a.tmpl = tmpl; a.emb_x = emb_x; a.emb_t = emb_t;
    a.out = nullptr; a.out2 = nullptr; a.feat_fused = nullptr;
    a.attn_w = const_cast<float*>(attn_w);
    a.B = B; a.N = N; a.CL = CL; a.E = E;
    a.alpha = alpha;
    a.beta = (float)(1.0 - (double)alpha);
    a.chunk_len = 0; a.n_chunks = 0;
    a.out_split = nullptr; a.split_c = 0; a.status = nullptr;
    a.g_feat = nullptr; a.g_s = nullptr;
    float* g_s = reinterpret_cast<float*>(ws);

    int rc = POF_ERR_UNSUPPORTED;
    switch (W) {
        case 1: rc = launch_gate_bwd_small<1>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 3: rc = launch_gate_bwd_small<3>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 5: rc = launch_gate_bwd_small<5>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 7: rc = launch_gate_bwd_small<7>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 9: rc = launch_gate_bwd_small<9>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 11: rc = launch_gate_bwd_small<11>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 13: rc = launch_gate_bwd_small<13>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
        case 15: rc = launch_gate_bwd_small<15>(a, g_out, g_feat, g_s, g_emb_x, g_emb_t, stream); break;
    }
    if (rc != POF_OK) return rc;

    // streaming part: g_tmpl (windowed, transposed weights) and g_x, one pass over g_out
    GateArgs s = a;
    s.tmpl = g_out;      // the register window marches over g_out rows
    s.x = g_out;
    s.out = g_tmpl;
    s.out2 = g_x;
    dim3 grid;
    plan_chunks(B, N, CL, s, grid);
    return dispatch_gate_stream<1>(W, s, grid, kGateThreads, stream);
}

}  // extern "C"
