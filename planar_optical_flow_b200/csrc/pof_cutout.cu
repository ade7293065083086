// Kernel 1 — distance-adaptive polar cutout.
//
// Replaces scans_to_cutout (/root/reference/src/utils/utils.py:259-334).  The
// arithmetic below is that function's, operation by operation and rounding by
// rounding (float32 half-angle and step, float64 sample angle / index / blend,
// float32 neighbour difference, float32 area means, float32-rounded clip
// bounds); see oracle/cutout.py for the same recipe in NumPy.
//
// Work decomposition (B200): the output [B, M, S, P] is a dense stream of
// "rows" of P floats, one row per (sample b, point m, scan s).  A CTA owns
// kTileRows consecutive rows.  Phase 1: one thread per row derives the row's
// geometry (range, half-angle, start angle, step, final clip values) once and
// parks it in shared memory; rows are partitioned into two-tap LINEAR rows and
// s_area-tap AREA rows so warps do not diverge between them.  Phase 2: every
// thread produces 16-byte pieces in address order (a warp writes up to 512
// contiguous bytes per store instruction) and the P samples of a row never
// recompute the arctangent.  The gathers from the range row hit L1 (a window
// spans a few 128-byte lines).
//
// The kernel is instruction-issue bound, not HBM bound, in EXACT arithmetic
// (~45 issue slots per 4-byte sample; tools/microbench.cu measured the FP64 and
// conversion rates this plan is built on: no __ddiv_rn, fmin/fmax(double) or
// floor(double) in the per-sample path).  POF_CUTOUT_FAST keeps the algorithm but
// evaluates the index line in 32.32 fixed point and the two-tap blend in float32
// (<= ~3e-6 of the output range from EXACT, inside the 1e-5 parity bar); it is
// what the streaming engine uses, EXACT is what `scans_to_cutout` uses.
//
// Area mode needs `s_area = ceil(max_span / P)` over a whole reference call
// (utils.py:308) = over one sample b here; cutout_span_kernel reduces it into
// `ws` first (one 8-byte slot per b, atomicMax on the bit pattern of a
// non-negative double).
#include <math.h>

#include "pof_common.cuh"

namespace pof {
namespace {

constexpr int kTileRows = 128;
constexpr int kThreads = 256;

struct CutoutArgs {
    const float* scans;
    const void* phi;
    float* out;
    unsigned long long* span_bits;  // [B]
    int* s_area_out;                // [B] or null
    const float* half_alpha_in;     // [B, S, M] or null: caller-supplied window half-angles
    float* half_alpha_out;          // [B, S, M] or null: the half-angles this call used
    int B, S, N, M, stride, P;
    long long rows;     // B*M*S
    float half_width;   // (float)(0.5 * window_width)      utils.py:279
    float depth_f;      // (float)window_depth              utils.py:327
    double depth;       // window_depth                     utils.py:330
    double pad;         // padding_val                      utils.py:326
    int fixed, centered, area_mode;
};

// 32.32 fixed-point view of a row's index line idx(k) = idx0 + k * slope, used by the FAST
// arithmetic: integer part = floor, fraction = blend ratio, one 64-bit multiply-add per sample.
struct FixedLine {
    long long base;    // idx(0) * 2^32
    long long slope;   // d idx / d k * 2^32 (round to nearest: error <= k * 2^-33 index units)
};
__device__ __forceinline__ long long to_fixed(double v) {
    return __double2ll_rn(v * 4294967296.0);
}

struct RowGeom {
    double start;   // phi[i] - half_alpha, evaluated in promote(phi, float)
    FixedLine lin;  // FAST only: index line of the P linear samples
    FixedLine are;  // FAST only: index line of the s_area*P area taps
    float step;     // 2*half_alpha/(P-1)
    float step_a;   // 2*half_alpha/(s_area*P-1) when the row is area-resampled
    int s_area;     // taps per sample; 0 = two-tap linear row
    float range;    // the point's reference range d
    float lo_f;     // final value of a sample clipped at d - window_depth
    float hi_f;     // final value of a sample clipped at d + window_depth
    float pad_f;    // final value of an out-of-scan sample
    int src;        // element offset of the (b, s) range row
};

// float32 arctangent.  NumPy's float32 arctan is a SIMD kernel that is within
// 1-2 ulp of the correctly rounded value; rounding the double result is the
// correctly rounded value in all but double-rounding cases, i.e. the closest
// any platform-independent code can get (SURVEY.md §7 hard part 1).
__device__ __forceinline__ float atan_f32(float x) { return (float)atan((double)x); }

// Correctly rounded a / b from y = RN(1/b) in three fused operations (Markstein):
// q0 = RN(a*y), r = a - b*q0 (exact in an fma), q = RN(q0 + r*y).  __ddiv_rn costs ~14
// DFMA-equivalents on sm_100 (tools/microbench.cu); b is a per-call constant here.
__device__ __forceinline__ double div_by(double a, double b, double y) {
    const double q0 = __dmul_rn(a, y);
    const double r = fma(-q0, b, a);
    return fma(r, y, q0);
}

struct Consts {
    double origin, pitch, inv_pitch, last;
    double depth, inv_depth;
};

// Fractional index of sample k on a row whose angular step is `step`  (:286-288).
// k*step is exact in double (k < 2^11, step has a 24-bit significand) so the
// fused form rounds exactly like NumPy's separate multiply and add.
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
__device__ __forceinline__ float finish(double v, float range, const Consts& c, int centered) {
    if (centered) v = div_by(__dsub_rn(v, (double)range), c.depth, c.inv_depth);
    return (float)v;
}

template <typename PhiT>
__device__ __forceinline__ void row_basics(const CutoutArgs& a, long long row, RowGeom& g, float& two_ha, int& b_out) {
    const int s = (int)(row % a.S);
    const long long bm = row / a.S;
    const int m = (int)(bm % a.M);
    const int b = (int)(bm / a.M);
    const int i = m * a.stride;
    const int src = (b * a.S + s) * a.N;
    const int ref = a.fixed ? src : (b * a.S + (a.S - 1)) * a.N;     // utils.py:274-278
    g.src = src;
    g.range = __ldg(a.scans + ref + i);
    const size_t ha_slot = ((size_t)b * a.S + s) * a.M + m;
    const float ha = a.half_alpha_in ? __ldg(a.half_alpha_in + ha_slot)
                                     : atan_f32(__fdiv_rn(a.half_width, fmaxf(g.range, 1e-2f)));   // :279
    if (a.half_alpha_out) a.half_alpha_out[ha_slot] = ha;
    two_ha = 2.0f * ha;
    g.step = __fdiv_rn(two_ha, (float)(a.P - 1));                    // :282
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    g.start = (double)(phi[i] - (PhiT)ha);                           // :284-285
    b_out = b;
}

template <typename PhiT>
__global__ void __launch_bounds__(kThreads) cutout_span_kernel(const CutoutArgs a) {
    const Consts c = make_consts<PhiT>(a);
    const long long row = (long long)blockIdx.x * kThreads + threadIdx.x;
    double span = 0.0;
    int b = -1;
    if (row < a.rows) {
        RowGeom g;
        float two_ha;
        row_basics<PhiT>(a, row, g, two_ha, b);
        span = __dsub_rn(sample_index(g.start, g.step, a.P - 1, c), sample_index(g.start, g.step, 0, c));   // :304
    }
    if (!(span > 0.0)) span = 0.0;
    // warp-level max when the whole warp belongs to one sample, else per-lane
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, b == b0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) span = fmax(span, __shfl_xor_sync(0xffffffffu, span, o));
        if ((threadIdx.x & 31) == 0 && b0 >= 0) atomicMax(a.span_bits + b0, (unsigned long long)__double_as_longlong(span));
    } else if (b >= 0) {
        atomicMax(a.span_bits + b, (unsigned long long)__double_as_longlong(span));
    }
}

// EXACT arithmetic: the reference's roundings, operation by operation (see file header).
__device__ __forceinline__ float sample_exact(const RowGeom& g, const float* __restrict__ src, int k, int s_area,
                                              const Consts& c, int nm1, int centered) {
    const double idx = sample_index(g.start, g.step, k, c);
    const int lo = __double2int_rd(idx);                         // floor; saturates far outside
    if (lo < 0 || (lo >= nm1 && idx > c.last)) return g.pad_f;   // :289, :326
    double v;
    if (s_area > 0) {                                            // :310-323
        float acc = 0.f;
        const int k0 = k * s_area;
        for (int t = 0; t < s_area; ++t) {
            const double ia = sample_index(g.start, g.step_a, k0 + t, c);
            // rint(clip(ia, 0, N-1)) == clip(rint(ia), 0, N-1): rint is monotone and fixes integers (:318)
            const int j = min(max(__double2int_rn(ia), 0), nm1);
            const float tap = __ldg(src + j);
            acc = (t == 0) ? tap : __fadd_rn(acc, tap);
        }
        v = (double)__fdiv_rn(acc, (float)s_area);
    } else {                                                     // :292-300
        const int hi = min(lo + 1, nm1);
        const double ratio = __dsub_rn(idx, (double)lo);
        const float v_lo = __ldg(src + lo);
        const float v_hi = __ldg(src + hi);
        v = __dadd_rn((double)v_lo, __dmul_rn(ratio, (double)__fsub_rn(v_hi, v_lo)));
    }
    return fminf(fmaxf(finish(v, g.range, c, centered), g.lo_f), g.hi_f);
}

// FAST arithmetic: same algorithm, index line in 32.32 fixed point and the blend / centring in
// float32.  Differs from EXACT by <= ~3e-6 of the output range (BASELINE tolerance 1e-5).
__device__ __forceinline__ float sample_fast(const RowGeom& g, const float* __restrict__ src, int k, int s_area, int nm1,
                                             float scale, float offset) {
    const long long fx = g.lin.base + (long long)k * g.lin.slope;
    const int lo = (int)(fx >> 32);
    const unsigned frac = (unsigned)fx;
    if (lo < 0 || lo > nm1 || (lo == nm1 && frac != 0u)) return g.pad_f;
    float v;
    if (s_area > 0) {
        float acc = 0.f;
        long long fa = g.are.base + (long long)(k * s_area) * g.are.slope + 0x80000000ll;   // +0.5: round to nearest
        for (int t = 0; t < s_area; ++t, fa += g.are.slope) {
            const int j = min(max((int)(fa >> 32), 0), nm1);
            const float tap = __ldg(src + j);
            acc = (t == 0) ? tap : acc + tap;
        }
        v = __fdiv_rn(acc, (float)s_area) - offset;
    } else {
        const float v_lo = __ldg(src + lo);
        const float v_hi = __ldg(src + min(lo + 1, nm1));
        v = fmaf((float)frac, (v_hi - v_lo) * 2.3283064365386963e-10f, v_lo - offset);
    }
    return fminf(fmaxf(v * scale, g.lo_f), g.hi_f);
}

// Tile of kTileRows output rows per CTA.
//   phase 1  one thread per row: geometry, clip constants, area decision; rows are split into a
//            LINEAR list and an AREA list (warp ballots) so that phase 2 has no divergence
//            between the 2-tap rows and the s_area-tap rows;
//   phase 2  16-byte pieces of the linear rows, then of the area rows, in address order.
template <typename PhiT, bool FAST>
__global__ void __launch_bounds__(kThreads) cutout_kernel(const CutoutArgs a) {
    __shared__ RowGeom geom[kTileRows];
    __shared__ unsigned char order[kTileRows];       // linear rows first, then area rows
    __shared__ int warp_lin[kTileRows / 32], warp_area[kTileRows / 32];
    const Consts c = make_consts<PhiT>(a);
    const long long row0 = (long long)blockIdx.x * kTileRows;
    const int rows_here = (int)min((long long)kTileRows, a.rows - row0);
    const double Pd = (double)a.P;
    const int tid = threadIdx.x;

    // ---- phase 1 ---------------------------------------------------------------------------------
    bool is_area = false;
    unsigned m_area = 0;
    if (tid < kTileRows) {
        if (tid < rows_here) {
            RowGeom g;
            float two_ha;
            int b;
            row_basics<PhiT>(a, row0 + tid, g, two_ha, b);
            g.step_a = 0.f;
            g.s_area = 0;
            g.are.base = g.are.slope = 0;
            double mx = 0.0;
            const double i0 = sample_index(g.start, g.step, 0, c);
            if (a.area_mode) {                                               // :304-310
                mx = __longlong_as_double((long long)a.span_bits[b]);
                const double span = __dsub_rn(sample_index(g.start, g.step, a.P - 1, c), i0);
                if (span > Pd) {
                    g.s_area = (int)ceil(__ddiv_rn(mx, Pd));                 // :308, one factor per sample b
                    g.step_a = __fdiv_rn(two_ha, (float)(g.s_area * a.P - 1));
                    is_area = true;
                }
            }
            // The depth clip, the centring and the float conversion are all monotone, so clipping the
            // FINAL float against the final values of the two bounds is the same function  (:327-334)
            g.lo_f = finish((double)(g.range - a.depth_f), g.range, c, a.centered);
            g.hi_f = finish((double)(g.range + a.depth_f), g.range, c, a.centered);
            g.pad_f = fminf(fmaxf(finish(a.pad, g.range, c, a.centered), g.lo_f), g.hi_f);      // :326
            if (FAST) {
                g.lin.base = to_fixed(i0);
                g.lin.slope = to_fixed((double)g.step * c.inv_pitch);
                if (is_area) {
                    g.are.base = g.lin.base;
                    g.are.slope = to_fixed((double)g.step_a * c.inv_pitch);
                }
            }
            geom[tid] = g;
            if (a.s_area_out && (row0 + tid) % ((long long)a.S * a.M) == 0)
                a.s_area_out[b] = (mx > Pd) ? (int)ceil(__ddiv_rn(mx, Pd)) : 0;
        }
        const bool valid = tid < rows_here;
        m_area = __ballot_sync(0xffffffffu, valid && is_area);
        const unsigned m_lin = __ballot_sync(0xffffffffu, valid && !is_area);
        if ((tid & 31) == 0) { warp_area[tid >> 5] = __popc(m_area); warp_lin[tid >> 5] = __popc(m_lin); }
        __syncwarp();
    }
    __syncthreads();
    int n_lin = 0;
#pragma unroll
    for (int w = 0; w < kTileRows / 32; ++w) n_lin += warp_lin[w];
    if (tid < rows_here) {
        const int w = tid >> 5;
        const unsigned lt = (1u << (tid & 31)) - 1u;
        int pos;
        if (is_area) {
            pos = n_lin + __popc(m_area & lt);
            for (int v = 0; v < w; ++v) pos += warp_area[v];
        } else {
            pos = (tid & 31) - __popc(m_area & lt);
            for (int v = 0; v < w; ++v) pos += warp_lin[v];
        }
        order[pos] = (unsigned char)tid;
    }
    __syncthreads();

    // ---- phase 2 ---------------------------------------------------------------------------------
    const unsigned vpr = (unsigned)a.P >> 2;                      // pieces per row
    const unsigned inv_vpr = 0xffffffffu / vpr + 1u;              // q / vpr == umulhi(q, inv_vpr) for q*vpr < 2^32
    float4* out4 = reinterpret_cast<float4*>(a.out + row0 * a.P);
    const float* scans = a.scans;
    const int nm1 = a.N - 1;
    const unsigned all_pieces = (unsigned)rows_here * vpr;
    const float scale = a.centered ? (float)c.inv_depth : 1.0f;

    for (unsigned q = tid; q < all_pieces; q += kThreads) {
        const unsigned li = __umulhi(q, inv_vpr);
        const unsigned cq = q - li * vpr;
        const unsigned r = order[li];
        const RowGeom& g = geom[r];
        const float* src = scans + g.src;
        const int s_area = g.s_area;          // warp-uniform except where the linear/area lists meet
        const float offset = a.centered ? g.range : 0.f;
        float res[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = (int)(cq << 2) + u;
            res[u] = FAST ? sample_fast(g, src, k, s_area, nm1, scale, offset)
                          : sample_exact(g, src, k, s_area, c, nm1, a.centered);
        }
        st_stream_f4(out4 + r * vpr + cq, make_float4(res[0], res[1], res[2], res[3]));
    }
}

}  // namespace
}  // namespace pof

extern "C" {

size_t pof_cutout_ws_bytes(int B) { return B > 0 ? (size_t)B * sizeof(unsigned long long) : 0; }

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
    POF_REQUIRE(numerics == POF_CUTOUT_EXACT || numerics == POF_CUTOUT_FAST, POF_ERR_BAD_PARAM,
                "pof_cutout_fwd: numerics must be POF_CUTOUT_EXACT or POF_CUTOUT_FAST");
    POF_REQUIRE(window_depth > 0.0, POF_ERR_BAD_PARAM, "pof_cutout_fwd: window_depth must be positive");
    POF_REQUIRE((long long)B * S * N < (1ll << 31), POF_ERR_BAD_SHAPE, "pof_cutout_fwd: B*S*N must fit int32");
    POF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, POF_ERR_BAD_PARAM, "pof_cutout_fwd: out must be 16-byte aligned");
    POF_REQUIRE(ws && ws_bytes >= pof_cutout_ws_bytes(B), POF_ERR_WORKSPACE,
                "pof_cutout_fwd: workspace too small (%zu < %zu)", ws_bytes, pof_cutout_ws_bytes(B));

    CutoutArgs a;
    a.scans = scans;
    a.phi = phi;
    a.out = out;
    a.span_bits = reinterpret_cast<unsigned long long*>(ws);
    a.s_area_out = s_area_out;
    a.half_alpha_in = half_alpha_in;
    a.half_alpha_out = half_alpha_out;
    a.B = B; a.S = S; a.N = N; a.stride = stride; a.P = P;
    a.M = (N + stride - 1) / stride;
    a.rows = (long long)B * a.M * S;
    a.half_width = (float)(0.5 * window_width);
    a.depth_f = (float)window_depth;
    a.depth = window_depth;
    a.pad = padding_val;
    a.fixed = fixed; a.centered = centered; a.area_mode = area_mode;

    if (area_mode) {
        POF_CUDA(cudaMemsetAsync(ws, 0, pof_cutout_ws_bytes(B), stream));
        const unsigned grid = (unsigned)((a.rows + kThreads - 1) / kThreads);
        if (phi_is_f64) cutout_span_kernel<double><<<grid, kThreads, 0, stream>>>(a);
        else cutout_span_kernel<float><<<grid, kThreads, 0, stream>>>(a);
        POF_CUDA(cudaGetLastError());
    }
    const unsigned grid = (unsigned)((a.rows + kTileRows - 1) / kTileRows);
    if (numerics == POF_CUTOUT_FAST) {
        if (phi_is_f64) cutout_kernel<double, true><<<grid, kThreads, 0, stream>>>(a);
        else cutout_kernel<float, true><<<grid, kThreads, 0, stream>>>(a);
    } else {
        if (phi_is_f64) cutout_kernel<double, false><<<grid, kThreads, 0, stream>>>(a);
        else cutout_kernel<float, false><<<grid, kThreads, 0, stream>>>(a);
    }
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
