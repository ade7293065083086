// Kernel 1 — distance-adaptive polar cutout.
//
// Replaces scans_to_cutout (/root/reference/src/utils/utils.py:259-334).  In EXACT
// arithmetic the code below is that function's, operation by operation and rounding by
// rounding (float32 half-angle and step, float64 sample angle / index / blend, float32
// neighbour difference, float32 area means, float32-rounded clip bounds); see
// oracle/cutout.py for the same recipe in NumPy.
//
// Work decomposition (B200).  The output [B, M, S, P] is made of "rows" of P floats, one per
// (sample b, point m, scan s).  A CTA owns one scan (b, s) and kTilePts consecutive points:
//   phase 0  the scan's N ranges are staged in shared memory (4.4 KB for a JRDB scan), so the
//            2 (linear) or s_area (area mode) gathers per sample are LDS with 32-bit addresses
//            instead of L1-wavefront-bound global gathers;
//   phase 1  one thread per point derives the row geometry once (range, half-angle via one
//            double arctangent, start angle, step, area decision, FINAL clip values) and parks
//            it in shared memory; rows are partitioned into two-tap LINEAR rows and s_area-tap
//            AREA rows (warp ballots) so warps do not diverge between the two;
//   phase 2  every thread produces 16-byte pieces (4 consecutive samples of a row) and stores
//            them with one streaming 128-bit store; a row is 4*P contiguous bytes.
//
// Two arithmetic policies, one algorithm (`numerics` argument):
//   POF_CUTOUT_EXACT  reproduces every rounding of the reference.  Issue bound (~45 slots per
//            4-byte sample); built on measured sm_100 rates (tools/microbench.cu): no __ddiv_rn
//            (14 DFMA-equivalents; replaced by Markstein's 3-operation correctly rounded
//            division by a constant), no fmin/fmax(double) (8), no floor(double) (4) in the
//            per-sample path; the clip is applied to the final float (all steps are monotone).
//   POF_CUTOUT_FAST   index line in 32.32 fixed point (one 64-bit add per sample), blend and
//            centring in float32: <= ~3e-6 of the output range from EXACT (parity bar 1e-5).
//            What the configs[1] cutout sweep measures and `StreamingDetector(cutout_fast=True)` selects;
//            `scans_to_cutout` and the engine's default use EXACT.
//
// Area mode needs `s_area = ceil(max_span / P)` over a whole reference call (utils.py:308) =
// over one sample b here: cutout_span_kernel (one CTA per b) reduces it into `ws` first.
#include <math_constants.h>
#include <math.h>
#include <stdlib.h>

#include "pof_common.cuh"

namespace pof {
namespace {

constexpr int kTilePts = 128;
constexpr int kThreads = 128;
constexpr int kMaxStagedPts = 8192;   // scans longer than this are gathered from global memory

struct CutoutArgs {
    const float* scans;
    const void* phi;
    float* out;
    double* span_max;               // [B]  max over the sample of idx[P-1] - idx[0]
    int* s_area_out;                // [B] or null
    const float* half_alpha_in;     // [B, S, M] or null: caller-supplied window half-angles
    float* half_alpha_out;          // [B, S, M] or null: the half-angles this call used
    int B, S, N, M, stride, P;
    int tiles_per_scan;
    float half_width;   // (float)(0.5 * window_width)      utils.py:279
    float depth_f;      // (float)window_depth              utils.py:327
    double depth;       // window_depth                     utils.py:330
    double pad;         // padding_val                      utils.py:326
    int fixed, centered, area_mode;
};

struct RowGeom {
    double start;          // phi[i] - half_alpha, evaluated in promote(phi, float)
    long long fx_base;     // FAST: idx(0) in 32.32 fixed point
    long long fx_slope;    // FAST: d idx / d sample, linear samples
    long long fx_slope_a;  // FAST: d idx / d tap, area taps
    float step;            // 2*half_alpha/(P-1)
    float step_a;          // 2*half_alpha/(s_area*P-1) when the row is area-resampled
    int s_area;            // taps per sample; 0 = two-tap linear row
    float range;           // the point's reference range d
    float lo_f;            // final value of a sample clipped at d - window_depth
    float hi_f;            // final value of a sample clipped at d + window_depth
    float pad_f;           // final value of an out-of-scan sample
    int pad_;
};

struct Consts {
    double origin, pitch, inv_pitch, last;
    double depth, inv_depth;
};

// float32 arctangent.  NumPy's float32 arctan is a SIMD kernel that is within 1-2 ulp of the
// correctly rounded value; rounding the double result is the correctly rounded value in all
// but double-rounding cases, i.e. the closest any platform-independent code can get
// (SURVEY.md §7 hard part 1).
__device__ __forceinline__ float atan_f32(float x) { return (float)atan((double)x); }

// Correctly rounded a / b from y = RN(1/b) in three fused operations (Markstein):
// q0 = RN(a*y), r = a - b*q0 (exact in an fma), q = RN(q0 + r*y).
__device__ __forceinline__ double div_by(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    const double r = fma(-q0, b, a);
    return fma(r, y, q0);
}

__device__ __forceinline__ long long to_fixed(double v) { return __double2ll_rn(v * 4294967296.0); }

// Fractional index of sample k on a row whose angular step is `step`  (:286-288).  k*step is
// exact in double (k < 2^11, step has a 24-bit significand) so the fused form rounds exactly
// like NumPy's separate multiply and add.
__device__ __forceinline__ double sample_index(double start, float step, int k, const Consts& c) {
    const double ang = fma((double)k, (double)step, start);
    return div_by(__dsub_rn(ang, c.origin), c.pitch, c.inv_pitch);
}

template <typename PhiT>
__device__ __forceinline__ Consts make_consts(const CutoutArgs& a) {
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    Consts c;
    c.origin = (double)phi[0];
    c.pitch = (double)(PhiT)(phi[1] - phi[0]);
    c.inv_pitch = __drcp_rn(c.pitch);
    c.last = (double)(a.N - 1);
    c.depth = a.depth;
    c.inv_depth = __drcp_rn(a.depth);
    return c;
}

// (v - d) / window_depth, or v itself when not centred, rounded to float  (:328-334)
// An infinite range (no return) is in the reference's domain: (inf - d) / depth = inf, where the three-operation division
// would produce inf - inf.
__device__ __forceinline__ double div_depth(double x, const Consts& c) {
    const double q = div_by(x, c.depth, c.inv_depth);
    return fabs(x) == CUDART_INF ? x : q;
}
__device__ __forceinline__ float finish(double v, float range, const Consts& c, int centered) {
    if (centered) v = div_depth(__dsub_rn(v, (double)range), c);
    return (float)v;
}
// np.clip = minimum(maximum(x, lo), hi) propagates a NaN from ANY of its operands (an infinite reference range makes
// both bounds NaN after centring; inf - inf inside a blend makes the sample NaN): max.NaN / min.NaN, not fmaxf / fminf.
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fmin_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float clip_nan(float v, float lo, float hi) { return fmin_nan(fmax_nan(v, lo), hi); }

// Range, half-angle, step and start angle of row (b, s, m)   (:274-285)
// FAST_ATAN (float32 arctangent, 1-2 ulp) was measured: it saves 9 % of the FAST kernel's instructions but the
// extra half-angle ulps against NumPy cost end-to-end parity of the streaming engine, so nothing uses it.
template <typename PhiT, bool FAST_ATAN = false>
__device__ __forceinline__ void row_basics(const CutoutArgs& a, int b, int s, int m, RowGeom& g, float& two_ha) {
    const int i = m * a.stride;
    const int ref_scan = a.fixed ? s : a.S - 1;
    g.range = __ldg(a.scans + ((size_t)b * a.S + ref_scan) * a.N + i);
    const size_t ha_slot = ((size_t)b * a.S + s) * a.M + m;
    const float ratio = __fdiv_rn(a.half_width, fmaxf(g.range, 1e-2f));
    const float ha = a.half_alpha_in ? __ldg(a.half_alpha_in + ha_slot) : (FAST_ATAN ? atanf(ratio) : atan_f32(ratio));   // :279
    if (a.half_alpha_out) a.half_alpha_out[ha_slot] = ha;
    two_ha = 2.0f * ha;
    g.step = __fdiv_rn(two_ha, (float)(a.P - 1));                    // :282
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    g.start = (double)(phi[i] - (PhiT)ha);                           // :284-285
}

// One CTA per sample b: max over its S*M rows of idx[P-1] - idx[0]   (:304, :308)
template <typename PhiT, bool FAST_ATAN>
__global__ void __launch_bounds__(kThreads) cutout_span_kernel(const CutoutArgs a) {
    __shared__ double warp_max[kThreads / 32];
    const Consts c = make_consts<PhiT>(a);
    const int b = blockIdx.x;
    double best = 0.0;
    for (int r = threadIdx.x; r < a.S * a.M; r += kThreads) {
        RowGeom g;
        float two_ha;
        row_basics<PhiT, FAST_ATAN>(a, b, r / a.M, r % a.M, g, two_ha);
        const double span = __dsub_rn(sample_index(g.start, g.step, a.P - 1, c), sample_index(g.start, g.step, 0, c));
        if (span > best) best = span;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other > best) best = other;
    }
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w)
            if (warp_max[w] > best) best = warp_max[w];
        a.span_max[b] = best;
        if (a.s_area_out) a.s_area_out[b] = (a.area_mode && best > (double)a.P) ? (int)ceil(__ddiv_rn(best, (double)a.P)) : 0;
    }
}

// The same reduction with one CTA per SCAN (b, s) and an atomic max per sample (non-negative doubles order like
// their bit patterns; `span_max` is zeroed first): B*S CTAs instead of B, and - the half-angle being a monotone
// function of the range - only the rows within 0.1 % of the scan's smallest reference range are evaluated (every other
// row has a strictly smaller float32 half-angle, hence a smaller span), as in cutout_scan_kernel.  Used when the caller
// wants neither `s_area_out` nor the half-angle tables from this pass and supplies no half-angles of its own.
template <typename PhiT>
__global__ void __launch_bounds__(kThreads) cutout_span_scan_kernel(const CutoutArgs a) {
    __shared__ double warp_max[kThreads / 32];
    __shared__ float warp_min[kThreads / 32];
    const Consts c = make_consts<PhiT>(a);
    const int b = blockIdx.x / a.S, sc = blockIdx.x - b * a.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* ref = a.scans + ((size_t)b * a.S + (a.fixed ? sc : a.S - 1)) * a.N;
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    float dmin = 3.0e38f;
    for (int m = threadIdx.x; m < a.M; m += kThreads) dmin = fminf(dmin, fmaxf(__ldg(ref + m * a.stride), 1e-2f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    if (lane == 0) warp_min[warp] = dmin;
    __syncthreads();
    dmin = warp_min[0];
    for (int w = 1; w < kThreads / 32; ++w) dmin = fminf(dmin, warp_min[w]);
    const float near = dmin * 1.001f;
    double best = 0.0;
    for (int m = threadIdx.x; m < a.M; m += kThreads) {
        const int i = m * a.stride;
        const float dc = fmaxf(__ldg(ref + i), 1e-2f);
        if (dc <= near) {
            const float ha = atan_f32(__fdiv_rn(a.half_width, dc));
            const float step = __fdiv_rn(2.0f * ha, (float)(a.P - 1));
            const double start = (double)(phi[i] - (PhiT)ha);
            const double span = __dsub_rn(sample_index(start, step, a.P - 1, c), sample_index(start, step, 0, c));
            if (span > best) best = span;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        if (other > best) best = other;
    }
    if (lane == 0) warp_max[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w)
            if (warp_max[w] > best) best = warp_max[w];
        if (best > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(a.span_max + b), (unsigned long long)__double_as_longlong(best));
    }
}

// ---- per-sample arithmetic --------------------------------------------------------------------
// EXACT: the reference's roundings, operation by operation.  `in_scan` (:289) is evaluated by the
// caller from the same idx.
__device__ __forceinline__ double index_exact(const RowGeom& g, float step, int k, const Consts& c) {
    return sample_index(g.start, step, k, c);
}
__device__ __forceinline__ bool in_scan(double idx, int lo, int nm1, const Consts& c) {
    return !(lo < 0 || (lo >= nm1 && idx > c.last));                 // :289
}
__device__ __forceinline__ float clip_final(float v, const RowGeom& g) { return fminf(fmaxf(v, g.lo_f), g.hi_f); }

__device__ __forceinline__ float linear_exact(const RowGeom& g, const float* src, int k, const Consts& c, int nm1, int centered) {
    const double idx = index_exact(g, g.step, k, c);
    const int lo = __double2int_rd(idx);                             // floor; saturates far outside
    if (!in_scan(idx, lo, nm1, c)) return g.pad_f;                   // :326
    const double ratio = __dsub_rn(idx, (double)lo);                 // :292-300
    const float v_lo = src[lo];
    const float v_hi = src[min(lo + 1, nm1)];
    const double v = __dadd_rn((double)v_lo, __dmul_rn(ratio, (double)__fsub_rn(v_hi, v_lo)));
    return clip_nan(finish(v, g.range, c, centered), g.lo_f, g.hi_f);
}

__device__ __forceinline__ float area_exact(const RowGeom& g, const float* src, int k, const Consts& c, int nm1, int centered) {
    const double idx = index_exact(g, g.step, k, c);
    const int lo = __double2int_rd(idx);
    if (!in_scan(idx, lo, nm1, c)) return g.pad_f;
    float acc = 0.f;                                                 // :310-323
    const int k0 = k * g.s_area;
    for (int t = 0; t < g.s_area; ++t) {
        const double ia = index_exact(g, g.step_a, k0 + t, c);
        // rint(clip(ia, 0, N-1)) == clip(rint(ia), 0, N-1): rint is monotone and fixes integers (:318)
        const int j = min(max(__double2int_rn(ia), 0), nm1);
        acc = (t == 0) ? src[j] : __fadd_rn(acc, src[j]);
    }
    return clip_nan(finish((double)__fdiv_rn(acc, (float)g.s_area), g.range, c, centered), g.lo_f, g.hi_f);
}

// FAST: `fx` is the sample's index in 32.32 fixed point; out = clip(v*scale + bias) with
// bias = -d*scale when centred.  The blend is one fused multiply-add on the raw fraction.
__device__ __forceinline__ float linear_fast(const RowGeom& g, const float* src, long long fx, int nm1, float scale,
                                             float bias, float dscale) {
    const int lo = (int)(fx >> 32);
    const unsigned frac = (unsigned)fx;
    if ((unsigned)lo < (unsigned)nm1) {                              // interior: beams lo and lo+1 exist
        const float v_lo = src[lo];
        const float t = fmaf(v_lo, scale, bias);
        return clip_final(fmaf((float)frac, (src[lo + 1] - v_lo) * dscale, t), g);
    }
    if (lo == nm1 && frac == 0u) return clip_final(fmaf(src[nm1], scale, bias), g);   // exactly the last beam
    return g.pad_f;
}

__device__ __forceinline__ float area_fast(const RowGeom& g, const float* src, long long fx, int k, int nm1, float scale,
                                           float bias) {
    const int lo = (int)(fx >> 32);
    if (!((unsigned)lo < (unsigned)nm1 || (lo == nm1 && (unsigned)fx == 0u))) return g.pad_f;
    float acc = 0.f;
    long long fa = g.fx_base + (long long)(k * g.s_area) * g.fx_slope_a + 0x80000000ll;   // +0.5: nearest tap
    for (int t = 0; t < g.s_area; ++t, fa += g.fx_slope_a) acc += src[min(max((int)(fa >> 32), 0), nm1)];
    return clip_final(fmaf(__fdiv_rn(acc, (float)g.s_area), scale, bias), g);
}

template <typename PhiT, bool FAST, bool STAGED>
__global__ void __launch_bounds__(kThreads) cutout_kernel(const CutoutArgs a) {
    extern __shared__ __align__(16) float staged[];       // STAGED: the scan's N ranges
    __shared__ RowGeom geom[kTilePts];
    __shared__ unsigned char order[kTilePts];              // linear rows first, then area rows
    __shared__ int warp_lin[kTilePts / 32], warp_area[kTilePts / 32];
    const Consts c = make_consts<PhiT>(a);
    const int tid = threadIdx.x;
    const int bs = blockIdx.x / a.tiles_per_scan;          // b * S + s
    const int m0 = (blockIdx.x - bs * a.tiles_per_scan) * kTilePts;
    const int b = bs / a.S, s = bs - b * a.S;
    const int rows_here = min(kTilePts, a.M - m0);
    const double Pd = (double)a.P;
    const float* scan = a.scans + (size_t)bs * a.N;

    // ---- phase 0: stage the scan ---------------------------------------------------------------
    if (STAGED)
        for (int i = tid; i < a.N; i += kThreads) staged[i] = __ldg(scan + i);
    const float* src = STAGED ? staged : scan;

    // ---- phase 1: one thread per point ---------------------------------------------------------
    static_assert(kThreads == kTilePts, "phase 1 maps one thread to one point");
    bool is_area = false;
    const bool valid = tid < rows_here;
    if (valid) {
        RowGeom g;
        float two_ha;
        row_basics<PhiT>(a, b, s, m0 + tid, g, two_ha);
        g.step_a = 0.f;
        g.s_area = 0;
        g.fx_base = g.fx_slope = g.fx_slope_a = 0;
        const double i0 = sample_index(g.start, g.step, 0, c);
        if (a.area_mode) {                                               // :304-310
            const double span = __dsub_rn(sample_index(g.start, g.step, a.P - 1, c), i0);
            if (span > Pd) {
                g.s_area = (int)ceil(__ddiv_rn(a.span_max[b], Pd));      // :308, one factor per sample b
                g.step_a = __fdiv_rn(two_ha, (float)(g.s_area * a.P - 1));
                is_area = true;
            }
        }
        // The depth clip, the centring and the float conversion are all monotone, so clipping the
        // FINAL float against the final values of the two bounds is the same function  (:327-334)
        g.lo_f = finish((double)(g.range - a.depth_f), g.range, c, a.centered);
        g.hi_f = finish((double)(g.range + a.depth_f), g.range, c, a.centered);
        g.pad_f = FAST ? fminf(fmaxf(finish(a.pad, g.range, c, a.centered), g.lo_f), g.hi_f)      // :326
                       : clip_nan(finish(a.pad, g.range, c, a.centered), g.lo_f, g.hi_f);
        if (FAST) {
            g.fx_base = to_fixed(i0);
            g.fx_slope = to_fixed((double)g.step * c.inv_pitch);
            g.fx_slope_a = to_fixed((double)g.step_a * c.inv_pitch);
        }
        geom[tid] = g;
    }
    const unsigned m_area = __ballot_sync(0xffffffffu, valid && is_area);
    const unsigned m_lin = __ballot_sync(0xffffffffu, valid && !is_area);
    if ((tid & 31) == 0) { warp_area[tid >> 5] = __popc(m_area); warp_lin[tid >> 5] = __popc(m_lin); }
    __syncthreads();
    int n_lin = 0;
#pragma unroll
    for (int w = 0; w < kTilePts / 32; ++w) n_lin += warp_lin[w];
    if (valid) {
        const int w = tid >> 5;
        const unsigned lt = (1u << (tid & 31)) - 1u;
        int pos;
        if (is_area) {
            pos = n_lin + __popc(m_area & lt);
            for (int v = 0; v < w; ++v) pos += warp_area[v];
        } else {
            pos = __popc(m_lin & lt);
            for (int v = 0; v < w; ++v) pos += warp_lin[v];
        }
        order[pos] = (unsigned char)tid;
    }
    __syncthreads();

    // ---- phase 2: 16-byte pieces, linear rows then area rows -----------------------------------
    // (a warp-per-point mapping with lanes along the samples was measured too: conflict-free
    //  gathers, but twice the instructions per sample; profiles/cutout_r1_notes.md)
    const unsigned vpr = (unsigned)a.P >> 2;                      // pieces per row
    const unsigned inv_vpr = 0xffffffffu / vpr + 1u;              // q / vpr == umulhi(q, inv_vpr) for q*vpr < 2^32
    const unsigned lin_pieces = (unsigned)n_lin * vpr;
    const unsigned pieces = (unsigned)rows_here * vpr;
    const int nm1 = a.N - 1;
    const float scale = a.centered ? (float)c.inv_depth : 1.0f;
    const float dscale = scale * 2.3283064365386963e-10f;         // scale * 2^-32
    // row (b, m, s) starts at ((b*M + m)*S + s)*P floats
    float4* out4 = reinterpret_cast<float4*>(a.out + (((size_t)b * a.M + m0) * a.S + s) * a.P);
    const size_t row_pitch4 = (size_t)a.S * vpr;

    for (unsigned q = tid; q < pieces; q += kThreads) {
        const unsigned li = vpr == 1u ? q : __umulhi(q, inv_vpr);        // (the magic number overflows for vpr = 1)
        const unsigned cq = q - li * vpr;
        const unsigned r = order[li];
        // by value: the store below goes through a generic pointer, so a reference into shared
        // memory would be reloaded after it
        const RowGeom g = geom[r];
        const int k0 = (int)(cq << 2);
        const bool area = q >= lin_pieces;                         // warp-uniform except in one warp
        float res[4];
        if (FAST) {
            const float bias = a.centered ? -g.range * scale : 0.f;
            long long fx = g.fx_base + (long long)k0 * g.fx_slope;
            if (!area) {
#pragma unroll
                for (int u = 0; u < 4; ++u, fx += g.fx_slope) res[u] = linear_fast(g, src, fx, nm1, scale, bias, dscale);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u, fx += g.fx_slope) res[u] = area_fast(g, src, fx, k0 + u, nm1, scale, bias);
            }
        } else {
            if (!area) {
#pragma unroll
                for (int u = 0; u < 4; ++u) res[u] = linear_exact(g, src, k0 + u, c, nm1, a.centered);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) res[u] = area_exact(g, src, k0 + u, c, nm1, a.centered);
            }
        }
        st_stream_f4(out4 + r * row_pitch4 + cq, make_float4(res[0], res[1], res[2], res[3]));
    }
}

// ---- FAST numerics, row-per-thread form ---------------------------------------------------------
// The piece-per-thread kernel above spends ~39 issue slots per 4-byte sample, most of them on per-piece
// bookkeeping (row lookup, 80 bytes of geometry from shared memory, 64-bit index set-up) and on address
// arithmetic for the scattered 16-byte stores.  Here a thread keeps ONE row's geometry in registers and
// walks its P samples with a running 32.32 fixed-point index (17 slots per sample, branch-free: the
// in-scan test is one 64-bit compare, the loads are clamped, the fraction becomes a float with one funnel
// shift); results are parked in a padded shared-memory tile and leave the SM as one TMA bulk store per row
// (cp.async.bulk.global.shared::cta, SASS UBLKCP), so the LSU never sees the output.  The few area-mode rows
// of a tile are done afterwards, spread over all threads as before.
constexpr int kRowPitchPad = 4;        // floats of padding per tile row: 16-byte aligned and conflict-free STS.128

struct AreaRow {
    long long fx_base, fx_slope, fx_slope_a;
    float lo_f, hi_f, pad_f, bias;
    int s_area, row;
};

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}

template <typename PhiT>
__global__ void __launch_bounds__(kThreads) cutout_rows_kernel(const CutoutArgs a) {
    extern __shared__ __align__(16) float smem_f[];          // [N + 1 (+pad)] staged scan | [kTilePts][P + pad] tile
    __shared__ AreaRow area_rows[kTilePts];
    __shared__ int warp_area[kTilePts / 32];
    const Consts c = make_consts<PhiT>(a);
    const int tid = threadIdx.x;
    const int bs = blockIdx.x / a.tiles_per_scan;            // b * S + s
    const int m0 = (blockIdx.x - bs * a.tiles_per_scan) * kTilePts;
    const int b = bs / a.S, s = bs - b * a.S;
    const int rows_here = min(kTilePts, a.M - m0);
    const int nm1 = a.N - 1;
    // S == 1: the tile's rows are adjacent in the output, so an unpadded tile leaves as ONE bulk store (its
    // 16-byte stores then have 2-way bank conflicts, which is far cheaper than issuing a store per row)
    const int pitch = a.S == 1 ? a.P : a.P + kRowPitchPad;
    float* staged = smem_f;
    float* tile = smem_f + ((a.N + 1 + 3) & ~3);
    const float* scan = a.scans + (size_t)bs * a.N;

    for (int i = tid; i < a.N; i += kThreads) staged[i] = __ldg(scan + i);
    if (tid == 0) staged[a.N] = __ldg(scan + nm1);           // beam N-1 once more: the two-tap load never leaves the array

    // ---- one thread per point: geometry in registers -------------------------------------------
    const bool valid = tid < rows_here;
    bool is_area = false;
    RowGeom g;
    float scale = a.centered ? (float)c.inv_depth : 1.0f, bias = 0.f;
    g.fx_base = g.fx_slope = g.fx_slope_a = 0;
    g.s_area = 0;
    g.lo_f = g.hi_f = g.pad_f = 0.f;
    if (valid) {
        float two_ha;
        row_basics<PhiT>(a, b, s, m0 + tid, g, two_ha);
        const double i0 = sample_index(g.start, g.step, 0, c);
        g.step_a = 0.f;
        if (a.area_mode) {                                               // :304-310
            const double span = __dsub_rn(sample_index(g.start, g.step, a.P - 1, c), i0);
            if (span > (double)a.P) {
                g.s_area = (int)ceil(__ddiv_rn(a.span_max[b], (double)a.P));
                g.step_a = __fdiv_rn(two_ha, (float)(g.s_area * a.P - 1));
                is_area = true;
            }
        }
        g.lo_f = finish((double)(g.range - a.depth_f), g.range, c, a.centered);
        g.hi_f = finish((double)(g.range + a.depth_f), g.range, c, a.centered);
        g.pad_f = fminf(fmaxf(finish(a.pad, g.range, c, a.centered), g.lo_f), g.hi_f);
        g.fx_base = to_fixed(i0);
        g.fx_slope = to_fixed((double)g.step * c.inv_pitch);
        g.fx_slope_a = to_fixed((double)g.step_a * c.inv_pitch);
        bias = a.centered ? -g.range * scale : 0.f;
    }
    const unsigned m_area = __ballot_sync(0xffffffffu, valid && is_area);
    if ((tid & 31) == 0) warp_area[tid >> 5] = __popc(m_area);
    __syncthreads();                                                     // staged scan + area counts visible

    // ---- two-tap rows: P samples per thread ------------------------------------------------------
    if (valid && !is_area) {
        // +2^-24 beam: the 23-bit fraction below is then rounded, not truncated; the in-scan limit moves with it
        long long fx = g.fx_base + 0x100ll;
        const unsigned long long limit = ((unsigned long long)(unsigned)nm1 << 32) + 0x100ull;
        float* dst = tile + tid * pitch;
        const long long step3 = 3 * g.fx_slope;
        for (int k = 0; k < a.P; k += 4) {
            float res[4];
            // the index grows along the row: if the first and the last sample of the chunk are inside the
            // scan, all four are, and neither clamping nor the padding select is needed
            if ((unsigned long long)fx <= limit && (unsigned long long)(fx + step3) <= limit) {
#pragma unroll
                for (int u = 0; u < 4; ++u, fx += g.fx_slope) {
                    const unsigned lo = (unsigned)((unsigned long long)fx >> 32);
                    const float v0 = staged[lo], v1 = staged[lo + 1];
                    const float w = __uint_as_float(__funnelshift_r((unsigned)fx, 0x7fu, 9)) - 1.0f;   // fraction of the index
                    res[u] = fminf(fmaxf(fmaf(fmaf(w, v1 - v0, v0), scale, bias), g.lo_f), g.hi_f);
                }
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u, fx += g.fx_slope) {
                    const unsigned lo = min((unsigned)((unsigned long long)fx >> 32), (unsigned)nm1);
                    const float v0 = staged[lo], v1 = staged[lo + 1];
                    const float w = __uint_as_float(__funnelshift_r((unsigned)fx, 0x7fu, 9)) - 1.0f;
                    const float v = fminf(fmaxf(fmaf(fmaf(w, v1 - v0, v0), scale, bias), g.lo_f), g.hi_f);
                    res[u] = ((unsigned long long)fx <= limit) ? v : g.pad_f;
                }
            }
            *reinterpret_cast<float4*>(dst + k) = make_float4(res[0], res[1], res[2], res[3]);
        }
    }

    // ---- area rows: compact them, then 4-sample pieces over all threads -------------------------
    int n_area = 0;
#pragma unroll
    for (int w = 0; w < kTilePts / 32; ++w) n_area += warp_area[w];
    if (n_area > 0) {                                                    // CTA-uniform
        if (valid && is_area) {
            int pos = __popc(m_area & ((1u << (tid & 31)) - 1u));
            for (int v = 0; v < (tid >> 5); ++v) pos += warp_area[v];
            AreaRow r;
            r.fx_base = g.fx_base; r.fx_slope = g.fx_slope; r.fx_slope_a = g.fx_slope_a;
            r.lo_f = g.lo_f; r.hi_f = g.hi_f; r.pad_f = g.pad_f; r.bias = bias; r.s_area = g.s_area; r.row = tid;
            area_rows[pos] = r;
        }
        __syncthreads();
        const unsigned vpr = (unsigned)a.P >> 2;
        const unsigned pieces = (unsigned)n_area * vpr;
        for (unsigned q = tid; q < pieces; q += kThreads) {
            const unsigned li = q / vpr, cq = q - li * vpr;
            const AreaRow r = area_rows[li];
            long long fx = r.fx_base + (long long)(cq << 2) * r.fx_slope;
            float res[4];
#pragma unroll
            for (int u = 0; u < 4; ++u, fx += r.fx_slope) {
                const int lo = (int)(fx >> 32);
                float v = r.pad_f;
                if ((unsigned)lo < (unsigned)nm1 || (lo == nm1 && (unsigned)fx == 0u)) {
                    float acc = 0.f;
                    long long fa = r.fx_base + (long long)(((int)(cq << 2) + u) * r.s_area) * r.fx_slope_a + 0x80000000ll;
                    for (int t = 0; t < r.s_area; ++t, fa += r.fx_slope_a) acc += staged[min(max((int)(fa >> 32), 0), nm1)];
                    v = fminf(fmaxf(fmaf(__fdiv_rn(acc, (float)r.s_area), scale, r.bias), r.lo_f), r.hi_f);
                }
                res[u] = v;
            }
            *reinterpret_cast<float4*>(tile + r.row * pitch + (cq << 2)) = make_float4(res[0], res[1], res[2], res[3]);
        }
    }

    // ---- the tile leaves through the TMA: one bulk store per row ---------------------------------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> async-proxy reads
    __syncthreads();
    if (a.S == 1 ? tid == 0 : valid) {
        float* row_out = a.out + (((size_t)b * a.M + m0 + tid) * a.S + s) * a.P;
        bulk_s2g(row_out, tile + tid * pitch, (unsigned)(a.S == 1 ? rows_here : 1) * (unsigned)a.P * 4u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory may go once it has been read
    }
}

// ---- FAST numerics, one CTA per scan (S == 1) -----------------------------------------------------
// The rows kernel above still spends half of its issue slots outside the sample loop: every 128-point
// tile re-stages the whole scan, the half-angles are computed twice (span kernel + tile kernel), and the
// loop itself carries an in-scan test per chunk, two 32-bit gathers, a fraction conversion and a
// separate scale step.  Here ONE CTA owns ONE scan and its warps work independently:
//   stage     the scan once, as (V, D) float2 pairs: V = v[i]*scale, D = (v[i+1]-v[i])*scale, so a two-tap
//             sample is ONE LDS.64, (V - d*scale) and ONE fma on the fraction w = w' - 1, w' = 1.fraction in
//             [1, 2) (the fraction bits dropped into a float's mantissa with one funnel shift; no int->float
//             conversion);
//   s_area    = ceil(max span / P) (utils.py:308) needs the scan's maximal index span.  The half-angle is
//             a monotone function of the range, so the maximum is attained by a row whose range is within
//             0.1 % of the scan's minimum: only those few rows pay a second arctangent, and no span kernel
//             runs.  (Caller-supplied half-angles are not monotone: then every row is examined.)
//   groups    a warp takes 32 consecutive rows at a time: a lane derives its row's geometry once (the double
//             arctangent included) and walks its P samples with a running 32.32 index, 9 issue slots per
//             sample (2 index adds, address, LDS.64, funnel shift, fma, bias add, max, min); rows that leave
//             the scan take a clamped loop.  Chunks are visited in an order rotated by one for every other
//             group of 4 lanes, which makes the unpadded tile's 16-byte stores bank-conflict free.  The
//             group's area rows are then resampled by the whole warp, two rows at a time.  The 32 rows are
//             adjacent in the output, so they leave as ONE TMA bulk store issued by lane 0.  No CTA-wide
//             barrier after the set-up.
constexpr int kScanWarpsMax = 4;       // measured 4 > 5 > 6 > 8 > 12 warps per CTA (five 4-warp CTAs fit an SM)
// tuning knobs (build.build_variant): launch bounds of the scan kernel and the unroll factor of its sample loop
#ifndef POF_SCAN_LB_THREADS
#define POF_SCAN_LB_THREADS (kScanWarpsMax * 32)
#define POF_SCAN_LB_BLOCKS 5
#endif
#ifndef POF_SCAN_UNROLL
#define POF_SCAN_UNROLL 2
#endif
constexpr int kScanUnroll = POF_SCAN_UNROLL;

__device__ __forceinline__ unsigned hi32(long long v) {
    unsigned lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ float frac_one_two(long long fx) {        // 1.fraction of a 32.32 index, 23 bits
    return __uint_as_float(__funnelshift_r((unsigned)fx, 0x7fu, 9));
}
__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}

// Four consecutive two-tap samples of a row, in two halves so that the loads of the next chunk are in flight while
// this one is finished.  INSIDE: every sample is known to lie in the scan.
struct ChunkTaps {
    float2 cd[4];
    float w[4];
    bool ok[4];
};
template <bool INSIDE>
__device__ __forceinline__ void chunk_load(ChunkTaps& t, long long& fx, long long slope, const float2* pairs, int nm1,
                                           unsigned long long limit) {
#pragma unroll
    for (int u = 0; u < 4; ++u, fx += slope) {
        t.cd[u] = pairs[INSIDE ? hi32(fx) : min(hi32(fx), (unsigned)nm1)];
        t.w[u] = frac_one_two(fx);
        t.ok[u] = INSIDE || (unsigned long long)fx <= limit;
    }
}
template <bool INSIDE>
__device__ __forceinline__ void chunk_finish(const ChunkTaps& t, float bias, float lo_f, float hi_f, float pad_f, float* dst) {
    float res[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        // (v*scale - d*scale) FIRST: both products are exact when scale is a power of two (every config of the reference)
        // and the difference is exact for a sample inside the depth window (Sterbenz), so the blend rounds once, at the
        // magnitude of the window instead of the range: 1e-7 of the output range.  (An fma on (C, D) = (v*scale - D, D)
        // with the bias added last saves one instruction and measured 0.3 % faster, at 5-8e-6.)
        const float v = fminf(fmaxf(fmaf(t.w[u] - 1.0f, t.cd[u].y, t.cd[u].x + bias), lo_f), hi_f);
        res[u] = t.ok[u] ? v : pad_f;
    }
    *reinterpret_cast<float4*>(dst) = make_float4(res[0], res[1], res[2], res[3]);
}
// One two-tap row of P = 4*nchunks samples into `dst`; chunks in the order rot, rot+1, ..., nchunks-1, (0 if rot).
template <bool INSIDE>
__device__ __forceinline__ void scan_row(long long fx0, long long slope, unsigned rot, int nchunks, const float2* pairs, int nm1,
                                         unsigned long long limit, float bias, float lo_f, float hi_f, float pad_f, float* dst) {
    long long fx = fx0 + (rot ? 4 * slope : 0);
    float* p = dst + 4 * rot;
    ChunkTaps t0, t1;
    chunk_load<INSIDE>(t0, fx, slope, pairs, nm1, limit);
#pragma unroll(INSIDE ? kScanUnroll : 1)
    for (int j = 0; j < nchunks - 2; ++j, p += 4) {
        chunk_load<INSIDE>(t1, fx, slope, pairs, nm1, limit);
        chunk_finish<INSIDE>(t0, bias, lo_f, hi_f, pad_f, p);
        t0 = t1;
    }
    if (nchunks > 1) {
        if (rot) fx = fx0;                                    // the rotated order ends with chunk 0
        chunk_load<INSIDE>(t1, fx, slope, pairs, nm1, limit);
        chunk_finish<INSIDE>(t0, bias, lo_f, hi_f, pad_f, p);
        p = rot ? dst : p + 4;
        chunk_finish<INSIDE>(t1, bias, lo_f, hi_f, pad_f, p);
    } else {
        chunk_finish<INSIDE>(t0, bias, lo_f, hi_f, pad_f, dst);
    }
}

// Arctangent for the scan kernel: degree-7 Taylor expansions around j/32, j = 0..32 (|h| <= 1/64: truncation
// 4e-16), coefficient k of interval j at [k][j] so that the lanes of a warp read one bank-conflict-free row per
// coefficient; a_0 = atan(c), a_k = cos^k(t) sin(k (t + pi/2)) / k with t = atan(c).  19 dependent DFMAs of the
// library arctangent become 7; the double result rounds to the same float32 (2e6 random ratios: no mismatch).
constexpr int kAtanDeg = 8, kAtanCells = 33;
__device__ const double kAtanTab[kAtanDeg][kAtanCells] = {
    {0.0, 0.031239833430268277, 0.06241880999595735, 0.09347678115858947, 0.12435499454676144, 0.15499674192394097, 0.18534794999569476, 0.21535769969773805, 0.24497866312686414, 0.2741674511196588, 0.3028848683749714, 0.3310960767041321, 0.35877067027057225, 0.38588266939807375, 0.4124104415973873, 0.43833655985795783, 0.4636476090008061, 0.48833395105640554, 0.5123894603107377, 0.5358112379604637, 0.5585993153435624, 0.5807563535676704, 0.6022873461349642, 0.6231993299340659, 0.6435011087932844, 0.6632029927060933, 0.6823165548747481, 0.7008544078844502, 0.7188299996216245, 0.7362574289814281, 0.7531512809621944, 0.7695264804056583, 0.7853981633974483},
    {1.0, 0.9990243902439024, 0.9961089494163423, 0.9912875121006776, 0.9846153846153846, 0.9761677788369877, 0.9660377358490566, 0.9543336439888164, 0.9411764705882353, 0.9266968325791856, 0.911032028469751, 0.8943231441048035, 0.8767123287671234, 0.8583403185247277, 0.839344262295082, 0.8198558847077663, 0.7999999999999999, 0.7798933739527799, 0.7596439169139466, 0.7393501805054151, 0.7191011235955057, 0.6989761092150172, 0.6790450928381964, 0.6593689632968449, 0.6400000000000001, 0.6209824135839903, 0.6023529411764706, 0.5841414717626925, 0.5663716814159292, 0.5490616621983913, 0.5322245322245323, 0.5158690176322418, 0.5000000000000001},
    {6.123233995736766e-17, -0.031189054134443818, -0.06201456494420795, -0.09212352484188287, -0.12118343195266268, -0.14889117694367784, -0.17498042007831957, -0.1992271540107127, -0.22145328719723184, -0.24152822423783293, -0.259368549030534, -0.27493602334051587, -0.2882341902796021, -0.2993039166020844, -0.3082182209083579, -0.31507672110466595, -0.31999999999999995, -0.3231241397032081, -0.3245956202837042, -0.3245667218392002, -0.3231915162226992, -0.32062248832251977, -0.3170077886989988, -0.31248909013939336, -0.30720000000000003, -0.3012649671723422, -0.29479861591695505, -0.28790543730916507, -0.28067977132116845, -0.2732060174370548, -0.2655590181577707, -0.2578045669980775, -0.25000000000000006},
    {-0.3333333333333333, -0.3313849680261942, -0.3255965744411859, -0.3161351980806101, -0.3032693066302534, -0.2873547671703596, -0.26881688015386307, -0.24812989172132635, -0.22579550851482458, -0.20232188302838702, -0.1782043459056766, -0.15390887298395117, -0.12985893504225618, -0.1064260321160807, -0.08392389965092538, -0.06260611727242969, -0.04266666666666668, -0.024242875586053887, -0.00742014386405217, 0.007762137828011928, 0.021304010058125548, 0.03323815990324352, 0.043623288662702764, 0.05253789717845137, 0.06007466666666659, 0.0663355373845743, 0.07142752588370986, 0.07545927443534463, 0.07853829252307457, 0.08076882835302857, 0.08225029603740784, 0.08307617931032006, 0.08333333333333333},
    {-6.123233995736766e-17, 0.03109782847033422, 0.0612925373519341, 0.0897296352784626, 0.11564771541612685, 0.13841508743188694, 0.1575558835513322, 0.172764162601911, 0.18390584403922366, 0.19100950613396706, 0.1942480277031656, 0.1939136496819828, 0.19038925685573274, 0.1841185660860868, 0.1755775325168041, 0.16524874751139976, 0.1536, 0.14106759025610235, 0.12804448210964214, 0.11487299052889793, 0.10184143725436906, 0.089184060640724, 0.07708341842577798, 0.06567454968857642, 0.05505024000000008, 0.04526683949711949, 0.03635019943918298, 0.02830140572395086, 0.021102089472027713, 0.014719180622174942, 0.009109038982802103, 0.004220948861087882, 2.2962127484012877e-17},
    {0.2, 0.19708362491474402, 0.18849239252123212, 0.17468634435813643, 0.15638847103500478, 0.13452479177016757, 0.11015087145241911, 0.08437309375308281, 0.05827291058184042, 0.03284107195122972, 0.008926832687212512, -0.012795245362054977, -0.03184060763886972, -0.047913800045917374, -0.06089436996927585, -0.07081332236715301, -0.07782400000000005, -0.082170912261112, -0.08415939044671553, -0.08412816077114978, -0.08242613450013074, -0.07939401190816328, -0.07535073678326323, -0.0705844389824263, -0.06534725632000002, -0.05985331032467883, -0.0542790925861128, -0.04876556844601561, -0.04342139512902429, -0.03832676006858236, -0.03353745621775853, -0.02908891415646316, -0.025000000000000043},
    {6.123233995736766e-17, -0.03096641713613131, -0.060260921564058025, -0.08635551414720784, -0.10799203289571492, -0.12427516125727307, -0.13472304594573917, -0.13927271513270162, -0.13824402393905807, -0.13227107456173015, -0.12221319485042205, -0.10905837633522088, -0.09383083087154981, -0.07751162873674149, -0.06097797755826145, -0.04496329362748313, -0.030037333333333343, -0.016603590630686293, -0.004910003678043123, 0.004931341550207509, 0.01291956154107462, 0.019138776750677104, 0.023732924964874693, 0.026883292056905875, 0.028789702655999977, 0.029655701234811937, 0.029677611118976936, 0.0290370748743366, 0.02789652856550813, 0.026397012476140822, 0.024657740472706552, 0.022776911998077104, 0.020833333333333343},
    {-0.14285714285714285, -0.13897938054121464, -0.12768213756913727, -0.10993120326594874, -0.08720236389717116, -0.06130445652312329, -0.0341739304551756, -0.00767210514433281, 0.016587229433004867, 0.03736038077505396, 0.05383747901208601, 0.06564868844587586, 0.07282245349402097, 0.07571029899280447, 0.07489423672083208, 0.07109147681735407, 0.0650678857142857, 0.05756754546991764, 0.04926176605657244, 0.04071759826693892, 0.03238358507275448, 0.02458919992709272, 0.01755399983204303, 0.011402731316661785, 0.0061832189893486125, 0.0018846241885264662, -0.001545572720249459, -0.004186837409626081, -0.006133528134935132, -0.007485570232337259, -0.008341653876645061, -0.008794611007126214, -0.008928571428571435},
};
__device__ __forceinline__ double atan_unit(float x, const double* tab) {          // 0 <= x <= 1
    const int j = __float2int_rn(x * 32.0f);
    const double h = fma((double)j, -0.03125, (double)x);
    double r = tab[(kAtanDeg - 1) * kAtanCells + j];
#pragma unroll
    for (int k = kAtanDeg - 2; k >= 0; --k) r = fma(r, h, tab[k * kAtanCells + j]);
    return r;
}
__device__ __forceinline__ float atan_f32_tab(float x, const double* tab) {         // x >= 0 (or NaN)
    if (x <= 1.0f) return (float)atan_unit(x, tab);
    if (!(x <= 3.0e38f)) return x == x ? 1.57079637050628662109375f : x;            // +inf -> pi/2, NaN stays
    const double inv = __drcp_rn((double)x);
    // the reciprocal is not a float: expand around the nearest cell of its float rounding, with the exact offset
    const float invf = (float)inv;
    const int j = __float2int_rn(invf * 32.0f);
    const double h = fma((double)j, -0.03125, inv);
    double r = tab[(kAtanDeg - 1) * kAtanCells + j];
#pragma unroll
    for (int k = kAtanDeg - 2; k >= 0; --k) r = fma(r, h, tab[k * kAtanCells + j]);
    return (float)(1.5707963267948966 - r);
}

// Sum of the s_area nearest-beam taps of an area sample, accumulated in tap order (:318-321).  TAPS > 0: unrolled.
template <int TAPS, bool CLAMP>
__device__ __forceinline__ float tap_sum(const float* vals, long long fa, long long rsa, int nm1, int s_area) {
    float acc = 0.f;
    if (TAPS > 0) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t, fa += rsa) acc += vals[CLAMP ? min(max((int)hi32(fa), 0), nm1) : hi32(fa)];
    } else {
#pragma unroll 1
        for (int t = 0; t < s_area; ++t, fa += rsa) acc += vals[CLAMP ? min(max((int)hi32(fa), 0), nm1) : hi32(fa)];
    }
    return acc;
}
// acc / s_area, correctly rounded: Markstein's three-operation form for a small integer divisor (checked against
// float division on 9e7 operands for every divisor 2..32), the plain division otherwise.
__device__ __forceinline__ float div_taps(float acc, float taps_f, float taps_rcp, int s_area) {
    if (s_area > 32) return __fdiv_rn(acc, taps_f);
    const float q0 = __fmul_rn(acc, taps_rcp);
    const float q = fmaf(fmaf(-q0, taps_f, acc), taps_rcp, q0);
    return q == q ? q : q0;                                   // an infinite sum stays infinite
}
// with a single chunk per row there is nothing to rotate
__device__ __forceinline__ unsigned nchunks_rot(int P, int lane) { return P >= 8 ? ((unsigned)lane >> 2) & 1u : 0u; }

// MULTI = false: S == 1, one CTA per sample; the scan's own span reduction happens here and the 32 rows of a group
// leave as one bulk store.  MULTI = true: S > 1 (training samples), one CTA per scan (b, s); the oversampling factor
// is a property of the whole sample (utils.py:308) and comes from cutout_span_kernel through `span_max`; a row's
// reference range is the newest scan's unless `fixed` (:274-278); the rows of one scan are S*P floats apart in the
// output, so a finished group is written with coalesced 16-byte stores, 14 lanes per 224-byte row.
template <typename PhiT, bool MULTI>
__global__ void __launch_bounds__(POF_SCAN_LB_THREADS, POF_SCAN_LB_BLOCKS) cutout_scan_kernel(const CutoutArgs a) {
    extern __shared__ __align__(16) float smem_f[];          // pairs [N+1] float2 | arctangent table | ranges [N] | per-warp tiles [32][P] | (MULTI, !fixed) reference ranges [N]
    __shared__ double warp_span[kScanWarpsMax];
    __shared__ float warp_min[kScanWarpsMax];
    __shared__ int next_group;                               // row groups are handed out dynamically (groups with area rows cost ~6x the others): +0.3 to +4 % (A/B, one GPU)
    if (threadIdx.x == 0) next_group = 0;
    const Consts c = make_consts<PhiT>(a);
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int b = MULTI ? blockIdx.x / a.S : blockIdx.x;     // S == 1: scan == sample
    const int sc = MULTI ? blockIdx.x - b * a.S : 0;
    const int nm1 = a.N - 1;
    const int P = a.P;
    float2* pairs = reinterpret_cast<float2*>(smem_f);
    double* atab = reinterpret_cast<double*>(smem_f + 2 * ((a.N + 2) & ~1));
    float* vals = reinterpret_cast<float*>(atab + kAtanDeg * kAtanCells);
    float* tile = vals + ((a.N + 3) & ~3) + (size_t)warp * 32 * P;
    const bool other_ref = MULTI && !a.fixed && sc != a.S - 1;
    float* dvals_w = other_ref ? vals + ((a.N + 3) & ~3) + (size_t)nwarps * 32 * P : vals;      // reference ranges d of the rows
    const float* dvals = dvals_w;
    const float* scan = a.scans + ((size_t)b * a.S + sc) * a.N;
    const float scale = a.centered ? (float)c.inv_depth : 1.0f;
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    const size_t row0 = ((size_t)b * a.S + sc) * a.M;        // [B, S, M] half-angle tables
    const float* ha_in = a.half_alpha_in ? a.half_alpha_in + row0 : nullptr;
    float* ha_out = a.half_alpha_out ? a.half_alpha_out + row0 : nullptr;

    // ---- stage the scan as (C, D) pairs; entry N repeats beam N-1 so index N-1 + 0 reads in bounds ---
    float dmin = 3.0e38f;
    for (int i = tid; i < kAtanDeg * kAtanCells; i += T) atab[i] = (&kAtanTab[0][0])[i];
    for (int i = tid; i <= a.N; i += T) {
        const float r0 = __ldg(scan + min(i, nm1));
        const float v0 = fminf(r0, 1e6f), v1 = fminf(__ldg(scan + min(i + 1, nm1)), 1e6f);     // finite pairs: inf - inf has no blend
        const float D = (v1 - v0) * scale;
        pairs[i] = make_float2(v0 * scale, D);
        if (i < a.N) vals[i] = r0;
        if (!MULTI && a.stride == 1) dmin = fminf(dmin, fmaxf(r0, 1e-2f));     // every beam is a row (entry N repeats beam N-1)
    }
    if (other_ref) {
        const float* ref = a.scans + ((size_t)b * a.S + (a.S - 1)) * a.N;
        for (int i = tid; i < a.N; i += T) dvals_w[i] = __ldg(ref + i);
    }
    if (!MULTI && a.stride != 1)
        for (int m = tid; m < a.M; m += T) dmin = fminf(dmin, fmaxf(__ldg(scan + m * a.stride), 1e-2f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    if (lane == 0) warp_min[warp] = dmin;
    __syncthreads();

    // ---- the scan's maximal index span, from the rows nearest to the sensor -----------------------------
    int s_area = 0;
    if (MULTI) {                                             // the sample's span was reduced over all its scans by cutout_span_kernel
        const double best = a.span_max[b];
        if (a.area_mode && best > (double)P) s_area = (int)ceil(__ddiv_rn(best, (double)P));
    } else if (a.area_mode || a.s_area_out) {
        dmin = warp_min[0];
        for (int w = 1; w < nwarps; ++w) dmin = fminf(dmin, warp_min[w]);
        const float near = a.half_alpha_in ? 3.0e38f : dmin * 1.001f;
        double best = 0.0;
        for (int m = tid; m < a.M; m += T) {
            const int i = m * a.stride;
            const float dc = fmaxf(vals[i], 1e-2f);
            if (dc <= near) {
                const float ha = ha_in ? __ldg(ha_in + m) : atan_f32_tab(__fdiv_rn(a.half_width, dc), atab);
                const float step = __fdiv_rn(2.0f * ha, (float)(P - 1));
                const double start = (double)(phi[i] - (PhiT)ha);
                const double span = __dsub_rn(sample_index(start, step, P - 1, c), sample_index(start, step, 0, c));
                if (span > best) best = span;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, best, o);
            if (other > best) best = other;
        }
        if (lane == 0) warp_span[warp] = best;
        __syncthreads();
        best = lane < nwarps ? warp_span[lane] : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, best, o);
            if (other > best) best = other;
        }
        best = __shfl_sync(0xffffffffu, best, 0);
        if (a.area_mode && best > (double)P) s_area = (int)ceil(__ddiv_rn(best, (double)P));
        if (tid == 0) {
            a.span_max[b] = best;
            if (a.s_area_out) a.s_area_out[b] = s_area;
        }
    }

    const unsigned rot = nchunks_rot(P, lane);
    const int nchunks = P >> 2;
    const unsigned long long limit = ((unsigned long long)(unsigned)nm1 << 32) + 0x100ull;
    const double span_unit = (double)(P - 1) * c.inv_pitch;
    float* out_b = a.out + (MULTI ? ((size_t)b * a.M * a.S + sc) * P : (size_t)b * a.M * P);
    const float taps_f = (float)s_area, taps_rcp = s_area > 0 ? 1.0f / (float)s_area : 0.f;
    bool store_pending = false;

    for (;;) {
        int m0 = 0;
        if (lane == 0) m0 = atomicAdd(&next_group, 32);
        m0 = __shfl_sync(0xffffffffu, m0, 0);
        if (m0 >= a.M) break;
        const int rows_here = min(32, a.M - m0);
        const bool valid = lane < rows_here;
        bool is_area = false;
        long long fx_base = 0, fx_slope = 0, fx_slope_a = 0;
        float lo_f = 0.f, hi_f = 0.f, pad_f = 0.f, bias = 0.f;
        if (valid) {                                          // :274-285, from the staged scan
            const int m = m0 + lane, i = m * a.stride;
            const float d = dvals[i];
            const float ratio = __fdiv_rn(a.half_width, fmaxf(d, 1e-2f));
            const float ha = ha_in ? __ldg(ha_in + m) : atan_f32_tab(ratio, atab);             // :279
            if (ha_out) ha_out[m] = ha;
            const float two_ha = 2.0f * ha;
            const float step = __fdiv_rn(two_ha, (float)(P - 1));                              // :282
            const double start = (double)(phi[i] - (PhiT)ha);                                  // :284-285
            fx_base = to_fixed(__dsub_rn(start, c.origin) * c.inv_pitch);
            fx_slope = to_fixed((double)step * c.inv_pitch);
            if (s_area > 0) {                                 // :304-310; the exact span only where the decision is close
                double span = (double)step * span_unit;
                if (fabs(span - (double)P) < 1e-6 * (double)P)
                    span = __dsub_rn(sample_index(start, step, P - 1, c), sample_index(start, step, 0, c));
                if (span > (double)P) {
                    is_area = true;
                    fx_slope_a = to_fixed((double)__fdiv_rn(two_ha, (float)(s_area * P - 1)) * c.inv_pitch);
                }
            }
            // float32 forms of the final clip values (EXACT rounds them through double; <= 1 ulp apart)
            if (a.centered) {
                lo_f = ((d - a.depth_f) - d) * scale;
                hi_f = ((d + a.depth_f) - d) * scale;
                pad_f = fminf(fmaxf(((float)a.pad - d) * scale, lo_f), hi_f);
                bias = -d * scale;
            } else {
                lo_f = d - a.depth_f;
                hi_f = d + a.depth_f;
                pad_f = fminf(fmaxf((float)a.pad, lo_f), hi_f);
            }
        }
        if (!MULTI && store_pending) {                        // the previous group's tile must have been read by the TMA
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
        const long long fx0 = fx_base + 0x100ll;              // +2^-24 beam: the 23-bit fraction is rounded, not truncated
        const long long fx_last = fx0 + (long long)(P - 1) * fx_slope;
        const bool inside = (unsigned long long)fx0 <= limit && (unsigned long long)fx_last <= limit;

        // ---- two-tap rows ---------------------------------------------------------------------------
        if (valid && !is_area) {
            float* dst = tile + lane * P;
            if (inside) scan_row<true>(fx0, fx_slope, rot, nchunks, pairs, nm1, limit, bias, lo_f, hi_f, pad_f, dst);
            else scan_row<false>(fx0, fx_slope, rot, nchunks, pairs, nm1, limit, bias, lo_f, hi_f, pad_f, dst);
        }

        // ---- area rows of the group: one at a time, the lanes along its samples (:310-323) --------------
        unsigned m_area = __ballot_sync(0xffffffffu, valid && is_area);      // 0 when s_area == 0
        while (m_area) {
            const int src = __ffs(m_area) - 1;
            m_area &= m_area - 1;
            const long long rb = shfl_ll(fx_base, src), rs = shfl_ll(fx_slope, src), rsa = shfl_ll(fx_slope_a, src);
            const float r_lo = __shfl_sync(0xffffffffu, lo_f, src), r_hi = __shfl_sync(0xffffffffu, hi_f, src);
            const float r_pad = __shfl_sync(0xffffffffu, pad_f, src), r_bias = __shfl_sync(0xffffffffu, bias, src);
            const bool r_inside = __shfl_sync(0xffffffffu, (int)inside, src) != 0;
            float* row = tile + src * P;
            for (int k = lane; k < P; k += 32) {
                const long long fx = rb + (long long)k * rs;
                const long long fa = rb + (long long)(k * s_area) * rsa + 0x80000000ll;        // +0.5: nearest tap
                const int lo = (int)hi32(fx);
                float v = r_pad;
                if (r_inside || (unsigned)lo < (unsigned)nm1 || (lo == nm1 && (unsigned)fx == 0u)) {
                    float acc;
                    if (!r_inside) acc = tap_sum<0, true>(vals, fa, rsa, nm1, s_area);
                    else switch (s_area) {                                                    // warp-uniform
                        case 2: acc = tap_sum<2, false>(vals, fa, rsa, nm1, s_area); break;
                        case 3: acc = tap_sum<3, false>(vals, fa, rsa, nm1, s_area); break;
                        case 4: acc = tap_sum<4, false>(vals, fa, rsa, nm1, s_area); break;
                        case 5: acc = tap_sum<5, false>(vals, fa, rsa, nm1, s_area); break;
                        case 6: acc = tap_sum<6, false>(vals, fa, rsa, nm1, s_area); break;
                        case 7: acc = tap_sum<7, false>(vals, fa, rsa, nm1, s_area); break;
                        case 8: acc = tap_sum<8, false>(vals, fa, rsa, nm1, s_area); break;
                        default: acc = tap_sum<0, false>(vals, fa, rsa, nm1, s_area); break;
                    }
                    v = fminf(fmaxf(fmaf(div_taps(acc, taps_f, taps_rcp, s_area), scale, r_bias), r_lo), r_hi);
                }
                row[k] = v;
            }
        }

        if (MULTI) {
            // ---- S > 1: the rows of this scan are S*P floats apart; 16-byte pieces, consecutive lanes along a row ----
            __syncwarp();
            const int q = P >> 2;                                 // pieces per row
            for (int t = lane; t < rows_here * q; t += 32) {
                const int r = t / q, c4 = t - r * q;
                st_stream_f4(reinterpret_cast<float4*>(out_b + ((size_t)(m0 + r) * a.S) * P + 4 * c4),
                             *reinterpret_cast<const float4*>(tile + r * P + 4 * c4));
            }
            __syncwarp();                                         // the tile is free for the next group
        } else {
            // ---- the group's rows are adjacent in the output (S == 1): one bulk store ------------------------
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(out_b + (unsigned)(m0 * P), tile, (unsigned)(rows_here * P) * 4u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
        }
    }
    if (!MULTI && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- EXACT numerics, one CTA per scan (round 2) ---------------------------------------------------------------
// The piece-per-thread EXACT kernel above issues ~60 instructions per sample: 14 double operations, five 64-bit
// conversions (F2I/I2F/F2F run at a quarter of the double rate) and ~40 slots of per-piece bookkeeping.  This is the
// same arithmetic, operation for operation, in the scan kernel's shape - a lane owns a row and walks its samples, a
// finished 32-row group leaves as one bulk store - with the conversions taken out of the sample loop:
//   * the scan is staged as DOUBLE pairs (v[i], (double)(v[i+1] - v[i])): one LDS.128 per sample, no F2F;
//   * floor(idx) is the low word of RD(idx + 1.5 * 2^52) and its double is that sum minus the constant (both exact for
//     |idx| < 2^31): no F2I / I2F.  Rows whose first and last index lie inside the scan (the index is monotone along a
//     row) take the plain loop, groups at the two ends of the scan the same loop with the :289 test and a clamped gather;
//     area rows are done by the whole warp with their taps on a guarded fixed-point line (exact_area_sample_fx); only
//     geometry outside those guarantees (|idx| >= 1e9, NaN) falls back to linear_exact / area_exact;
//   * (double)k comes from a P-entry table in shared memory;
//   * for a power-of-two window_depth the correctly rounded division is one product.
// 12 double operations and one F2F.F32.F64 per sample remain.
constexpr double kFloorMagic = 6755399441055744.0;      // 1.5 * 2^52

struct ExactRow {
    double start, step_d, range_d;
    float lo_f, hi_f, pad_f;
};

// MODE 0: not centred; 1: centred, window_depth a power of two; 2: centred, any window_depth
template <int MODE>
__device__ __forceinline__ float exact_finish(double v, const ExactRow& r, const Consts& c) {
    if (MODE != 0) {
        v = __dsub_rn(v, r.range_d);                                                             // :329-330
        v = MODE == 1 ? __dmul_rn(v, c.inv_depth) : div_depth(v, c);
    }
    return clip_nan((float)v, r.lo_f, r.hi_f);
}
// INSIDE: every sample of the row lies in the scan.  Otherwise (rows at the two ends of the scan) the :289 test selects the
// padding value and the gather is clamped; |idx| < 2^30 is the caller's guarantee either way.
template <int MODE, bool INSIDE>
__device__ __forceinline__ float exact_sample(double kd, const ExactRow& r, const Consts& c, const double2* dpairs, int nm1) {
    const double ang = fma(kd, r.step_d, r.start);                                               // :286-288
    const double idx = div_by(__dsub_rn(ang, c.origin), c.pitch, c.inv_pitch);
    const double t = __dadd_rd(idx, kFloorMagic);
    const double ratio = __dsub_rn(idx, __dsub_rn(t, kFloorMagic));                              // :292-294
    const int lo = __double2loint(t);
    const double2 vd = dpairs[INSIDE ? lo : min(max(lo, 0), nm1)];
    const float res = exact_finish<MODE>(__dadd_rn(vd.x, __dmul_rn(ratio, vd.y)), r, c);         // :300
    return INSIDE || in_scan(idx, lo, nm1, c) ? res : r.pad_f;
}
// A lane keeps kExactChunks x 4 samples in flight.  All lanes walk the samples in the same order: (double)k is a broadcast
// load and the gathers of neighbouring rows stay on neighbouring entries (the rotated order of the FAST kernel cost this
// one 1.6x the ideal gather wavefronts, more than it saves in store conflicts).
#ifndef POF_EXACT_CHUNKS
#define POF_EXACT_CHUNKS 2
#endif
constexpr int kExactChunks = POF_EXACT_CHUNKS;
template <int MODE, bool INSIDE>
__device__ __forceinline__ void exact_row(const ExactRow& r, const Consts& c, const double2* dpairs, const double* ktab,
                                          int nm1, int nchunks, float* dst) {
    int ch = 0;
#pragma unroll 1
    for (; ch + kExactChunks <= nchunks; ch += kExactChunks) {
        float res[4 * kExactChunks];
#pragma unroll
        for (int u = 0; u < 4 * kExactChunks; ++u) res[u] = exact_sample<MODE, INSIDE>(ktab[4 * ch + u], r, c, dpairs, nm1);
#pragma unroll
        for (int q = 0; q < kExactChunks; ++q)
            *reinterpret_cast<float4*>(dst + 4 * (ch + q)) = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
    }
#pragma unroll 1
    for (; ch < nchunks; ++ch) {
        float res[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) res[u] = exact_sample<MODE, INSIDE>(ktab[4 * ch + u], r, c, dpairs, nm1);
        *reinterpret_cast<float4*>(dst + 4 * ch) = make_float4(res[0], res[1], res[2], res[3]);
    }
}
template <bool INSIDE>
__device__ __forceinline__ void exact_row_mode(int mode, const ExactRow& r, const Consts& c, const double2* dpairs, const double* ktab,
                                               int nm1, int nchunks, float* dst) {
    if (mode == 1) exact_row<1, INSIDE>(r, c, dpairs, ktab, nm1, nchunks, dst);
    else if (mode == 0) exact_row<0, INSIDE>(r, c, dpairs, ktab, nm1, nchunks, dst);
    else exact_row<2, INSIDE>(r, c, dpairs, ktab, nm1, nchunks, dst);
}
// One sample of an area-resampled row that lies inside the scan: the s_area nearest-beam taps in tap order (:310-323), the
// rounding to the nearest beam (ties to even, like np.rint) taken from RN(ia + 1.5 * 2^52) instead of an F2I.
template <int MODE>
__device__ __forceinline__ float exact_area_sample(int k, const ExactRow& r, double step_a_d, int s_area, float s_area_f,
                                                   const Consts& c, const float* vals, int nm1) {
    double kd = (double)(k * s_area);
    float acc = 0.f;
#pragma unroll 2
    for (int t = 0; t < s_area; ++t) {
        const double ia = div_by(__dsub_rn(fma(kd, step_a_d, r.start), c.origin), c.pitch, c.inv_pitch);
        const float tap = vals[min(max(__double2loint(__dadd_rn(ia, kFloorMagic)), 0), nm1)];
        acc = t == 0 ? tap : __fadd_rn(acc, tap);
        kd = __dadd_rn(kd, 1.0);
    }
    return exact_finish<MODE>((double)__fdiv_rn(acc, s_area_f), r, c);
}

// The same sample with the taps located on a 32.32 fixed-point index line (integer adds, no double operation per tap).
// The line is within 2^-20 beam of the reference's double index for every tap - |start| < 64 rad, 1 / pitch < 1e5 and
// s_area * P <= 4096 (the caller's conditions) bound the three roundings of the double index by 1.4e-9 beam, the
// rounding of the line's origin by 3e-12 and the accumulated quantisation of its slope by 4096 * 2^-33 = 2^-21 - so both
// round to the same beam unless the line passes within 2^-19 of a half-way point; then (one tap in 2.6e5, exact ties
// included) the sample is recomputed with the double index.
constexpr unsigned kTapGuard = 1u << 13;                      // 2^-19 beam in 32.32
template <int MODE>
__device__ __forceinline__ float exact_area_sample_fx(int k, const ExactRow& r, long long fx_base, long long fx_slope_a, double step_a_d,
                                                      int s_area, float s_area_f, const Consts& c, const float* vals, int nm1) {
    long long fa = fx_base + (long long)(k * s_area) * fx_slope_a + 0x80000000ll;      // +0.5: truncation rounds to nearest
    float acc = 0.f;
    bool close = false;
#pragma unroll 4
    for (int t = 0; t < s_area; ++t, fa += fx_slope_a) {
        close |= (unsigned)fa + kTapGuard < 2u * kTapGuard;
        const float tap = vals[min(max((int)hi32(fa), 0), nm1)];
        acc = t == 0 ? tap : __fadd_rn(acc, tap);
    }
    if (close) return exact_area_sample<MODE>(k, r, step_a_d, s_area, s_area_f, c, vals, nm1);
    return exact_finish<MODE>((double)__fdiv_rn(acc, s_area_f), r, c);
}

template <typename PhiT, bool MULTI>
__global__ void __launch_bounds__(kScanWarpsMax * 32, 4) cutout_scan_exact_kernel(const CutoutArgs a, const int mode) {
    extern __shared__ __align__(16) float smem_f[];          // pairs [N+1] double2 | (double)k [P] | ranges [N] | per-warp tiles [32][P] | (MULTI, !fixed) reference ranges [N]
    __shared__ double warp_span[kScanWarpsMax];
    __shared__ float warp_min[kScanWarpsMax];
    __shared__ int next_group;                               // row groups are handed out dynamically: groups with area rows cost several times the others
    if (threadIdx.x == 0) next_group = 0;
    const Consts c = make_consts<PhiT>(a);
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int b = MULTI ? blockIdx.x / a.S : blockIdx.x;
    const int sc = MULTI ? blockIdx.x - b * a.S : 0;
    const int nm1 = a.N - 1;
    const int P = a.P;
    double2* dpairs = reinterpret_cast<double2*>(smem_f);
    double* ktab = reinterpret_cast<double*>(dpairs + (a.N + 1));
    float* vals = reinterpret_cast<float*>(ktab + P);
    float* tile = vals + ((a.N + 3) & ~3) + (size_t)warp * 32 * P;
    const bool other_ref = MULTI && !a.fixed && sc != a.S - 1;
    float* dvals_w = other_ref ? vals + ((a.N + 3) & ~3) + (size_t)nwarps * 32 * P : vals;
    const float* dvals = dvals_w;
    const float* scan = a.scans + ((size_t)b * a.S + sc) * a.N;
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    const size_t row0 = ((size_t)b * a.S + sc) * a.M;
    const float* ha_in = a.half_alpha_in ? a.half_alpha_in + row0 : nullptr;
    float* ha_out = a.half_alpha_out ? a.half_alpha_out + row0 : nullptr;
    const double Pd = (double)P;

    // ---- stage the scan; entry N repeats beam N-1 (inds_ct_high is clipped at N-1, :293) -----------------
    float dmin = 3.0e38f;
    for (int k = tid; k < P; k += T) ktab[k] = (double)k;
#pragma unroll 4
    for (int i = tid; i <= a.N; i += T) {                    // (unrolled: the loads of four rounds are in flight together)
        const float v0 = __ldg(scan + min(i, nm1)), v1 = __ldg(scan + min(i + 1, nm1));
        dpairs[i] = make_double2((double)v0, (double)__fsub_rn(v1, v0));
        if (i < a.N) vals[i] = v0;
        if (!MULTI && a.stride == 1) dmin = fminf(dmin, fmaxf(v0, 1e-2f));
    }
    if (other_ref) {
        const float* ref = a.scans + ((size_t)b * a.S + (a.S - 1)) * a.N;
        for (int i = tid; i < a.N; i += T) dvals_w[i] = __ldg(ref + i);
    }
    if (!MULTI && a.stride != 1)
        for (int m = tid; m < a.M; m += T) dmin = fminf(dmin, fmaxf(__ldg(scan + m * a.stride), 1e-2f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    if (lane == 0) warp_min[warp] = dmin;
    __syncthreads();

    // ---- the sample's oversampling factor (:304-308) ---------------------------------------------------------
    int s_area = 0;
    if (MULTI) {                                             // reduced over all scans of the sample by the span kernels
        const double best = a.span_max[b];
        if (a.area_mode && best > Pd) s_area = (int)ceil(__ddiv_rn(best, Pd));
    } else if (a.area_mode || a.s_area_out) {                // S == 1: from the rows nearest to the sensor (see cutout_span_scan_kernel)
        dmin = warp_min[0];
        for (int w = 1; w < nwarps; ++w) dmin = fminf(dmin, warp_min[w]);
        const float near = ha_in ? 3.0e38f : dmin * 1.001f;
        double best = 0.0;
        for (int m = tid; m < a.M; m += T) {
            const int i = m * a.stride;
            const float dc = fmaxf(vals[i], 1e-2f);
            if (dc <= near) {
                const float ha = ha_in ? __ldg(ha_in + m) : atan_f32(__fdiv_rn(a.half_width, dc));
                const float step = __fdiv_rn(2.0f * ha, (float)(P - 1));
                const double start = (double)(phi[i] - (PhiT)ha);
                const double span = __dsub_rn(sample_index(start, step, P - 1, c), sample_index(start, step, 0, c));
                if (span > best) best = span;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, best, o);
            if (other > best) best = other;
        }
        if (lane == 0) warp_span[warp] = best;
        __syncthreads();
        best = lane < nwarps ? warp_span[lane] : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, best, o);
            if (other > best) best = other;
        }
        best = __shfl_sync(0xffffffffu, best, 0);
        if (a.area_mode && best > Pd) s_area = (int)ceil(__ddiv_rn(best, Pd));
        if (tid == 0) {
            a.span_max[b] = best;
            if (a.s_area_out) a.s_area_out[b] = s_area;
        }
    }

    const int nchunks = P >> 2;
    float* out_b = a.out + (MULTI ? ((size_t)b * a.M * a.S + sc) * P : (size_t)b * a.M * P);
    bool store_pending = false;
    const bool fx_taps = c.inv_pitch > 0.0 && c.inv_pitch < 1.0e5 && (long long)s_area * P <= 4096;

    for (;;) {
        int m0 = 0;
        if (lane == 0) m0 = atomicAdd(&next_group, 32);
        m0 = __shfl_sync(0xffffffffu, m0, 0);
        if (m0 >= a.M) break;
        const int rows_here = min(32, a.M - m0);
        const bool valid = lane < rows_here;
        bool is_area = false, inside = false, bounded = false;
        RowGeom g;
        g.start = 0.0; g.step = g.step_a = g.range = g.lo_f = g.hi_f = g.pad_f = 0.f; g.s_area = 0;
        if (valid) {                                          // :274-285, as in cutout_kernel's phase 1
            const int m = m0 + lane, i = m * a.stride;
            const PhiT phi_i = phi[i];                       // (issued ahead of the arctangent that hides its latency)
            g.range = dvals[i];
            const float ratio = __fdiv_rn(a.half_width, fmaxf(g.range, 1e-2f));
            const float ha = ha_in ? __ldg(ha_in + m) : atan_f32(ratio);                       // :279
            if (ha_out) ha_out[m] = ha;
            const float two_ha = 2.0f * ha;
            g.step = __fdiv_rn(two_ha, (float)(P - 1));                                        // :282
            g.start = (double)(phi_i - (PhiT)ha);                                              // :284-285
            const double i0 = sample_index(g.start, g.step, 0, c);
            const double i1 = sample_index(g.start, g.step, P - 1, c);
            if (s_area > 0 && __dsub_rn(i1, i0) > Pd) {                                        // :304-310
                is_area = true;
                g.s_area = s_area;
                g.step_a = __fdiv_rn(two_ha, (float)(s_area * P - 1));
            }
            inside = g.step >= 0.f && i0 >= 0.0 && i1 <= c.last;     // the index is monotone in k: every sample lies in the scan
            bounded = g.step >= 0.f && fabs(i0) < 1.0e9 && fabs(i1) < 1.0e9;    // floor(idx) fits the low word of idx + 1.5 * 2^52
            if (mode == 1) {                                 // :327-334 (see cutout_kernel); the division is one exact product
                const double rd = (double)g.range;
                g.lo_f = (float)__dmul_rn(__dsub_rn((double)(g.range - a.depth_f), rd), c.inv_depth);
                g.hi_f = (float)__dmul_rn(__dsub_rn((double)(g.range + a.depth_f), rd), c.inv_depth);
                g.pad_f = clip_nan((float)__dmul_rn(__dsub_rn(a.pad, rd), c.inv_depth), g.lo_f, g.hi_f);      // :326
            } else {
                g.lo_f = finish((double)(g.range - a.depth_f), g.range, c, a.centered);
                g.hi_f = finish((double)(g.range + a.depth_f), g.range, c, a.centered);
                g.pad_f = clip_nan(finish(a.pad, g.range, c, a.centered), g.lo_f, g.hi_f);
            }
        }
        if (!MULTI && store_pending) {                        // the previous group's tile must have been read by the TMA
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }

        // ---- two-tap rows ---------------------------------------------------------------------------
        ExactRow r;
        r.start = g.start; r.step_d = (double)g.step; r.range_d = (double)g.range; r.lo_f = g.lo_f; r.hi_f = g.hi_f; r.pad_f = g.pad_f;
        const bool two_tap = valid && !is_area;
        if (__all_sync(0xffffffffu, !two_tap || inside)) {            // warp-uniform: the common group
            if (two_tap) exact_row_mode<true>(mode, r, c, dpairs, ktab, nm1, nchunks, tile + lane * P);
        } else if (two_tap) {                                         // a group at one end of the scan
            if (bounded) exact_row_mode<false>(mode, r, c, dpairs, ktab, nm1, nchunks, tile + lane * P);
            else {
                float* dst = tile + lane * P;
#pragma unroll 1
                for (int k = 0; k < P; ++k) dst[k] = linear_exact(g, vals, k, c, nm1, a.centered);
            }
        }

        // ---- area rows of the group: one at a time, the lanes along its samples (:310-323) --------------
        unsigned m_area = __ballot_sync(0xffffffffu, valid && is_area);
        while (m_area) {
            const int src = __ffs(m_area) - 1;
            m_area &= m_area - 1;
            RowGeom ra;
            ra.start = __shfl_sync(0xffffffffu, g.start, src);
            ra.step = __shfl_sync(0xffffffffu, g.step, src);
            ra.step_a = __shfl_sync(0xffffffffu, g.step_a, src);
            ra.range = __shfl_sync(0xffffffffu, g.range, src);
            ra.lo_f = __shfl_sync(0xffffffffu, g.lo_f, src);
            ra.hi_f = __shfl_sync(0xffffffffu, g.hi_f, src);
            ra.pad_f = __shfl_sync(0xffffffffu, g.pad_f, src);
            ra.s_area = s_area;
            const bool ra_inside = __shfl_sync(0xffffffffu, (int)inside, src) != 0;
            float* row = tile + src * P;
            if (ra_inside) {
                ExactRow rr;
                rr.start = ra.start; rr.step_d = 0.0; rr.range_d = (double)ra.range; rr.lo_f = ra.lo_f; rr.hi_f = ra.hi_f; rr.pad_f = ra.pad_f;
                const double step_a_d = (double)ra.step_a;
                const float s_area_f = (float)s_area;
                if (fx_taps && fabs(ra.start) < 64.0) {       // (see exact_area_sample_fx)
                    const long long fx_base = to_fixed(__dsub_rn(ra.start, c.origin) * c.inv_pitch);
                    const long long fx_slope_a = to_fixed(step_a_d * c.inv_pitch);
                    for (int k = lane; k < P; k += 32)
                        row[k] = mode == 1 ? exact_area_sample_fx<1>(k, rr, fx_base, fx_slope_a, step_a_d, s_area, s_area_f, c, vals, nm1)
                               : mode == 0 ? exact_area_sample_fx<0>(k, rr, fx_base, fx_slope_a, step_a_d, s_area, s_area_f, c, vals, nm1)
                                           : exact_area_sample_fx<2>(k, rr, fx_base, fx_slope_a, step_a_d, s_area, s_area_f, c, vals, nm1);
                } else {
                    for (int k = lane; k < P; k += 32)
                        row[k] = mode == 1 ? exact_area_sample<1>(k, rr, step_a_d, s_area, s_area_f, c, vals, nm1)
                               : mode == 0 ? exact_area_sample<0>(k, rr, step_a_d, s_area, s_area_f, c, vals, nm1)
                                           : exact_area_sample<2>(k, rr, step_a_d, s_area, s_area_f, c, vals, nm1);
                }
            } else {
                for (int k = lane; k < P; k += 32) row[k] = area_exact(ra, vals, k, c, nm1, a.centered);
            }
        }

        if (MULTI) {
            __syncwarp();
            const int q = P >> 2;
            for (int t = lane; t < rows_here * q; t += 32) {
                const int r = t / q, c4 = t - r * q;
                st_stream_f4(reinterpret_cast<float4*>(out_b + ((size_t)(m0 + r) * a.S) * P + 4 * c4),
                             *reinterpret_cast<const float4*>(tile + r * P + 4 * c4));
            }
            __syncwarp();
        } else {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(out_b + (unsigned)(m0 * P), tile, (unsigned)(rows_here * P) * 4u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
        }
    }
    if (!MULTI && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <typename PhiT, bool MULTI>
bool launch_cutout_scan_exact(const CutoutArgs& a, cudaStream_t stream, int* status) {
    const int warps = kScanWarpsMax;
    const size_t smem = sizeof(double2) * (size_t)(a.N + 1) + sizeof(double) * (size_t)a.P +
                        ((size_t)((a.N + 3) & ~3) + (size_t)warps * 32 * a.P + (MULTI && !a.fixed ? (size_t)((a.N + 3) & ~3) : 0)) * sizeof(float);
    if (smem > 110 * 1024) return false;
    static bool attr_set[2][2][64] = {{{false}}};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    const int which = sizeof(PhiT) == 8;
    if (dev < 64 && !attr_set[which][MULTI][dev]) {
        if (cudaFuncSetAttribute(cutout_scan_exact_kernel<PhiT, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) != cudaSuccess) return false;
        attr_set[which][MULTI][dev] = true;
    }
    int depth_exp = 0;
    const int mode = !a.centered ? 0 : (frexp(a.depth, &depth_exp) == 0.5 ? 1 : 2);
    cutout_scan_exact_kernel<PhiT, MULTI><<<MULTI ? a.B * a.S : a.B, warps * 32, smem, stream>>>(a, mode);
    const cudaError_t e = cudaGetLastError();
    *status = e == cudaSuccess ? POF_OK : cuda_fail(e, "cutout_scan_exact_kernel launch");
    return true;
}

int scan_warps_for(int M) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("POF_SCAN_WARPS");           // tuning aid
        forced = e ? atoi(e) : 0;
    }
    if (forced > 0 && forced <= kScanWarpsMax) return forced;
    // small CTAs: five 4-warp CTAs fit an SM (43 KB each), and their set-up phases overlap the others' sample loops
    // (measured on JRDB- and DROW-shaped scans: 4 warps > 5 > 6 > 8 > 12)
    (void)M;
    return 4;
}

template <typename PhiT, bool MULTI>
bool launch_cutout_scan(const CutoutArgs& a, cudaStream_t stream, int* status) {
    const int warps = scan_warps_for(a.M);
    const size_t smem = ((size_t)2 * ((a.N + 2) & ~1) + (size_t)((a.N + 3) & ~3) + (size_t)warps * 32 * a.P +
                         (MULTI && !a.fixed ? (size_t)((a.N + 3) & ~3) : 0)) * sizeof(float) +
                        sizeof(double) * kAtanDeg * kAtanCells;
    if (smem > 110 * 1024) return false;
    static bool attr_set[2][2][64] = {{{false}}};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    const int which = sizeof(PhiT) == 8;
    if (dev < 64 && !attr_set[which][MULTI][dev]) {
        if (cudaFuncSetAttribute(cutout_scan_kernel<PhiT, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) != cudaSuccess) return false;
        attr_set[which][MULTI][dev] = true;
    }
    cutout_scan_kernel<PhiT, MULTI><<<MULTI ? a.B * a.S : a.B, warps * 32, smem, stream>>>(a);
    const cudaError_t e = cudaGetLastError();
    *status = e == cudaSuccess ? POF_OK : cuda_fail(e, "cutout_scan_kernel launch");
    return true;
}

template <typename PhiT, bool FAST>
int launch_cutout(const CutoutArgs& a, cudaStream_t stream) {
    const unsigned grid = (unsigned)a.tiles_per_scan * (unsigned)(a.B * a.S);
    if (FAST && a.S == 1 && a.N < kMaxStagedPts) {      // S > 1: rows are not adjacent, a bulk store per row is slower than pieces
        const size_t smem = ((size_t)((a.N + 1 + 3) & ~3) + (size_t)kTilePts * (a.P + kRowPitchPad)) * sizeof(float);
        if (smem <= 200 * 1024) {
            static bool attr_set[2][64] = {{false}};
            int dev = 0;
            POF_CUDA(cudaGetDevice(&dev));
            const int which = sizeof(PhiT) == 8;
            if (dev < 64 && !attr_set[which][dev]) {
                POF_CUDA(cudaFuncSetAttribute(cutout_rows_kernel<PhiT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_set[which][dev] = true;
            }
            cutout_rows_kernel<PhiT><<<grid, kThreads, smem, stream>>>(a);
            POF_CUDA(cudaGetLastError());
            return POF_OK;
        }
    }
    if (a.N <= kMaxStagedPts) {
        cutout_kernel<PhiT, FAST, true><<<grid, kThreads, (size_t)a.N * sizeof(float), stream>>>(a);
    } else {
        cutout_kernel<PhiT, FAST, false><<<grid, kThreads, 0, stream>>>(a);
    }
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // namespace
}  // namespace pof

extern "C" {

size_t pof_cutout_ws_bytes(int B) { return B > 0 ? (size_t)B * sizeof(double) : 0; }

int pof_cutout_fwd(const float* scans, const void* phi, int phi_is_f64, int B, int S, int N, int stride, int P,
                   double window_width, double window_depth, double padding_val, int fixed, int centered,
                   int area_mode, int numerics, float* out, int* s_area_out, const float* half_alpha_in,
                   float* half_alpha_out, void* ws, size_t ws_bytes, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(scans && phi && out, POF_ERR_NULL_POINTER, "pof_cutout_fwd: null scans/phi/out");
    POF_REQUIRE(B >= 0 && S >= 1 && N >= 2 && stride >= 1, POF_ERR_BAD_SHAPE,
                "pof_cutout_fwd: need B>=0, S>=1, N>=2, stride>=1 (got B=%d S=%d N=%d stride=%d)", B, S, N, stride);
    POF_REQUIRE(P >= 4 && (P % 4) == 0 && P <= 1024, POF_ERR_BAD_SHAPE,
                "pof_cutout_fwd: num_cutout_pts must be a multiple of 4 in [4,1024] (got %d)", P);
    POF_REQUIRE(numerics == POF_CUTOUT_EXACT || numerics == POF_CUTOUT_FAST || numerics == POF_CUTOUT_EXACT_PIECES, POF_ERR_BAD_PARAM,
                "pof_cutout_fwd: numerics must be POF_CUTOUT_EXACT, POF_CUTOUT_FAST or POF_CUTOUT_EXACT_PIECES");
    POF_REQUIRE(window_depth > 0.0, POF_ERR_BAD_PARAM, "pof_cutout_fwd: window_depth must be positive");
    POF_REQUIRE((long long)B * S * N < (1ll << 31), POF_ERR_BAD_SHAPE, "pof_cutout_fwd: B*S*N must fit int32");
    POF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, POF_ERR_BAD_PARAM, "pof_cutout_fwd: out must be 16-byte aligned");
    POF_REQUIRE(ws && ws_bytes >= pof_cutout_ws_bytes(B), POF_ERR_WORKSPACE,
                "pof_cutout_fwd: workspace too small (%zu < %zu)", ws_bytes, pof_cutout_ws_bytes(B));

    CutoutArgs a;
    a.scans = scans;
    a.phi = phi;
    a.out = out;
    a.span_max = reinterpret_cast<double*>(ws);
    a.s_area_out = s_area_out;
    a.half_alpha_in = half_alpha_in;
    a.half_alpha_out = half_alpha_out;
    a.B = B; a.S = S; a.N = N; a.stride = stride; a.P = P;
    a.M = (N + stride - 1) / stride;
    a.tiles_per_scan = (a.M + kTilePts - 1) / kTilePts;
    a.half_width = (float)(0.5 * window_width);
    a.depth_f = (float)window_depth;
    a.depth = window_depth;
    a.pad = padding_val;
    a.fixed = fixed; a.centered = centered; a.area_mode = area_mode;
    POF_REQUIRE((long long)a.tiles_per_scan * B * S < (1ll << 31), POF_ERR_BAD_SHAPE, "pof_cutout_fwd: too many tiles");

    // The scan kernel blends in output units around the row's own range: exact products when scale = 1 / window_depth is
    // a power of two (every config of the reference: 1e-7 of the output range).  For another scale v * scale rounds at
    // the magnitude of range * scale: still within the FAST contract (1e-5) while 1.5 ulp(padding_val * scale) <= 6e-6,
    // otherwise the rows kernel, which blends in metres first, takes the call.
    int mag_exp = 0, scale_exp = 0;
    frexp(centered ? fabs(padding_val / window_depth) : 0.0, &mag_exp);           // |x| = m * 2^e, 0.5 <= m < 1
    const bool scale_pow2 = frexp((double)(float)(1.0 / window_depth), &scale_exp) == 0.5;      // v * scale is then exact
    const bool coarse = !scale_pow2 && 1.5 * ldexp(1.0, mag_exp - 1 - 23) > 6.0e-6;             // 1.5 float32 ulps of padding_val * scale
    if (numerics == POF_CUTOUT_FAST && S == 1 && !coarse) {        // one CTA per scan: span reduction, half-angles and samples in one launch
        int status = POF_OK;
        if (phi_is_f64 ? launch_cutout_scan<double, false>(a, stream, &status) : launch_cutout_scan<float, false>(a, stream, &status)) return status;
    }
    if (numerics == POF_CUTOUT_EXACT && S == 1) {                  // the same shape with the reference's own roundings
        int status = POF_OK;
        if (phi_is_f64 ? launch_cutout_scan_exact<double, false>(a, stream, &status) : launch_cutout_scan_exact<float, false>(a, stream, &status)) return status;
    }
    if (area_mode || s_area_out) {
        if (!s_area_out && !half_alpha_in && !half_alpha_out) {       // one CTA per scan, nearest rows only, atomic max per sample
            POF_CUDA(cudaMemsetAsync(a.span_max, 0, (size_t)B * sizeof(double), stream));
            if (phi_is_f64) cutout_span_scan_kernel<double><<<B * S, kThreads, 0, stream>>>(a);
            else cutout_span_scan_kernel<float><<<B * S, kThreads, 0, stream>>>(a);
        } else if (phi_is_f64) {
            cutout_span_kernel<double, false><<<B, kThreads, 0, stream>>>(a);
        } else {
            cutout_span_kernel<float, false><<<B, kThreads, 0, stream>>>(a);
        }
        POF_CUDA(cudaGetLastError());
    }
    if (numerics == POF_CUTOUT_FAST && S > 1 && !coarse && N < kMaxStagedPts) {      // one CTA per scan (b, s) of the training samples
        int status = POF_OK;
        if (phi_is_f64 ? launch_cutout_scan<double, true>(a, stream, &status) : launch_cutout_scan<float, true>(a, stream, &status)) return status;
    }
    if (numerics == POF_CUTOUT_EXACT && S > 1) {
        int status = POF_OK;
        if (phi_is_f64 ? launch_cutout_scan_exact<double, true>(a, stream, &status) : launch_cutout_scan_exact<float, true>(a, stream, &status)) return status;
    }
    if (numerics == POF_CUTOUT_FAST) return phi_is_f64 ? launch_cutout<double, true>(a, stream) : launch_cutout<float, true>(a, stream);
    return phi_is_f64 ? launch_cutout<double, false>(a, stream) : launch_cutout<float, false>(a, stream);
}

}  // extern "C"
