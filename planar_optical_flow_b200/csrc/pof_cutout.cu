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
// geometry (range, half-angle, start angle, step) once and parks it in shared
// memory.  Phase 2: every thread produces 16-byte pieces of the tile in address
// order, so a warp writes 512 contiguous bytes per store instruction and the
// P samples of a row never recompute the arctangent.  The gathers from the
// range row hit L1 (a window spans at most a few 128-byte lines).
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

struct RowGeom {
    double start;   // phi[i] - half_alpha, evaluated in promote(phi, float)
    float step;     // 2*half_alpha/(P-1)
    float two_ha;   // 2*half_alpha
    float range;    // the point's reference range d
    int src;        // element offset of the (b, s) range row
};

// float32 arctangent.  NumPy's float32 arctan is a SIMD kernel that is within
// 1-2 ulp of the correctly rounded value; rounding the double result is the
// correctly rounded value in all but double-rounding cases, i.e. the closest
// any platform-independent code can get (SURVEY.md §7 hard part 1).
__device__ __forceinline__ float atan_f32(float x) { return (float)atan((double)x); }

template <typename PhiT>
__device__ __forceinline__ RowGeom row_geometry(const CutoutArgs& a, long long row) {
    const int s = (int)(row % a.S);
    const long long bm = row / a.S;
    const int m = (int)(bm % a.M);
    const int b = (int)(bm / a.M);
    const int i = m * a.stride;
    const int src = (b * a.S + s) * a.N;
    const int ref = a.fixed ? src : (b * a.S + (a.S - 1)) * a.N;     // utils.py:274-278
    RowGeom g;
    g.src = src;
    g.range = __ldg(a.scans + ref + i);
    const size_t ha_slot = ((size_t)b * a.S + s) * a.M + m;
    const float ha = a.half_alpha_in ? __ldg(a.half_alpha_in + ha_slot)
                                     : atan_f32(__fdiv_rn(a.half_width, fmaxf(g.range, 1e-2f)));   // :279
    if (a.half_alpha_out) a.half_alpha_out[ha_slot] = ha;
    g.two_ha = 2.0f * ha;
    g.step = __fdiv_rn(g.two_ha, (float)(a.P - 1));                  // :282
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    g.start = (double)(phi[i] - (PhiT)ha);                           // :284-285
    return g;
}

// Fractional index of sample k on a row whose angular step is `step`  (:286-288).
// k*step is exact in double (k < 2^11, step has a 24-bit significand) so the
// fused form rounds exactly like NumPy's separate multiply and add.
__device__ __forceinline__ double sample_index(double start, float step, int k, double origin, double pitch) {
    const double ang = fma((double)k, (double)step, start);
    return __ddiv_rn(__dsub_rn(ang, origin), pitch);
}

template <typename PhiT>
__global__ void __launch_bounds__(kThreads) cutout_span_kernel(const CutoutArgs a) {
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    const double origin = (double)phi[0];
    const double pitch = (double)(PhiT)(phi[1] - phi[0]);
    const long long row = (long long)blockIdx.x * kThreads + threadIdx.x;
    double span = 0.0;
    int b = -1;
    if (row < a.rows) {
        const RowGeom g = row_geometry<PhiT>(a, row);
        const double i0 = sample_index(g.start, g.step, 0, origin, pitch);
        const double i1 = sample_index(g.start, g.step, a.P - 1, origin, pitch);
        span = __dsub_rn(i1, i0);                                    // :304
        b = (int)(row / ((long long)a.S * a.M));
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

template <typename PhiT>
__global__ void __launch_bounds__(kThreads) cutout_kernel(const CutoutArgs a) {
    __shared__ RowGeom geom[kTileRows];
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    const double origin = (double)phi[0];
    const double pitch = (double)(PhiT)(phi[1] - phi[0]);
    const double last = (double)(a.N - 1);
    const long long row0 = (long long)blockIdx.x * kTileRows;
    const int rows_here = (int)min((long long)kTileRows, a.rows - row0);
    const long long rows_per_b = (long long)a.S * a.M;

    if (threadIdx.x < rows_here) geom[threadIdx.x] = row_geometry<PhiT>(a, row0 + threadIdx.x);
    __syncthreads();

    const int vpr = a.P >> 2;                       // 16-byte pieces per row
    const int pieces = rows_here * vpr;
    float4* out4 = reinterpret_cast<float4*>(a.out + row0 * a.P);
    const float* scans = a.scans;
    const int nm1 = a.N - 1;
    const double Pd = (double)a.P;

    for (int q = threadIdx.x; q < pieces; q += kThreads) {
        const int r = q / vpr;
        const int c0 = (q - r * vpr) << 2;
        const RowGeom g = geom[r];
        const float* src = scans + g.src;
        const double lo_b = (double)(g.range - a.depth_f);           // :327 bounds in float32
        const double hi_b = (double)(g.range + a.depth_f);
        const double dd = (double)g.range;

        // area-mode decision for this row (:304-308)
        int s_area = 0;
        float step_a = 0.f;
        if (a.area_mode) {
            const double i0 = sample_index(g.start, g.step, 0, origin, pitch);
            const double i1 = sample_index(g.start, g.step, a.P - 1, origin, pitch);
            if (__dsub_rn(i1, i0) > Pd) {
                const int b = (int)((row0 + r) / rows_per_b);
                const double mx = __longlong_as_double((long long)a.span_bits[b]);
                s_area = (int)ceil(__ddiv_rn(mx, Pd));               // :308
                step_a = __fdiv_rn(g.two_ha, (float)(s_area * a.P - 1));   // :310
            }
        }

        float res[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u;
            const double idx = sample_index(g.start, g.step, c, origin, pitch);
            double v;
            if (idx < 0.0 || idx > last) {                           // :289, :326
                v = a.pad;
            } else if (s_area > 0) {                                 // :310-323
                float acc = 0.f;
                for (int t = 0; t < s_area; ++t) {
                    double ia = sample_index(g.start, step_a, c * s_area + t, origin, pitch);
                    ia = fmin(fmax(ia, 0.0), last);
                    const int j = __double2int_rn(ia);               // rint: half to even (:318)
                    const float tap = __ldg(src + j);
                    acc = (t == 0) ? tap : __fadd_rn(acc, tap);
                }
                v = (double)__fdiv_rn(acc, (float)s_area);
            } else {                                                 // :292-300
                const double fl = floor(idx);
                const int lo = (int)fl;                              // 0 <= idx <= N-1 here
                const int hi = min(lo + 1, nm1);
                const double ratio = __dsub_rn(idx, fl);
                const float v_lo = __ldg(src + lo);
                const float v_hi = __ldg(src + hi);
                v = __dadd_rn((double)v_lo, __dmul_rn(ratio, (double)__fsub_rn(v_hi, v_lo)));
            }
            v = fmin(fmax(v, lo_b), hi_b);                           // :327
            if (a.centered) v = __ddiv_rn(__dsub_rn(v, dd), a.depth);   // :328-330
            res[u] = (float)v;
        }
        st_stream_f4(out4 + q, make_float4(res[0], res[1], res[2], res[3]));
    }

    // report the factor each sample used (first row of each sample does it)
    if (a.s_area_out && threadIdx.x < rows_here) {
        const long long row = row0 + threadIdx.x;
        if (row % rows_per_b == 0) {
            const int b = (int)(row / rows_per_b);
            const double mx = a.area_mode ? __longlong_as_double((long long)a.span_bits[b]) : 0.0;
            a.s_area_out[b] = (mx > Pd) ? (int)ceil(__ddiv_rn(mx, Pd)) : 0;
        }
    }
}

}  // namespace
}  // namespace pof

extern "C" {

size_t pof_cutout_ws_bytes(int B) { return B > 0 ? (size_t)B * sizeof(unsigned long long) : 0; }

int pof_cutout_fwd(const float* scans, const void* phi, int phi_is_f64, int B, int S, int N, int stride, int P,
                   double window_width, double window_depth, double padding_val, int fixed, int centered,
                   int area_mode, float* out, int* s_area_out, const float* half_alpha_in, float* half_alpha_out,
                   void* ws, size_t ws_bytes, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(scans && phi && out, POF_ERR_NULL_POINTER, "pof_cutout_fwd: null scans/phi/out");
    POF_REQUIRE(B >= 0 && S >= 1 && N >= 2 && stride >= 1, POF_ERR_BAD_SHAPE,
                "pof_cutout_fwd: need B>=0, S>=1, N>=2, stride>=1 (got B=%d S=%d N=%d stride=%d)", B, S, N, stride);
    POF_REQUIRE(P >= 4 && (P % 4) == 0 && P <= 1024, POF_ERR_BAD_SHAPE,
                "pof_cutout_fwd: num_cutout_pts must be a multiple of 4 in [4,1024] (got %d)", P);
    POF_REQUIRE(window_depth != 0.0, POF_ERR_BAD_PARAM, "pof_cutout_fwd: window_depth must be non-zero");
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
    if (phi_is_f64) cutout_kernel<double><<<grid, kThreads, 0, stream>>>(a);
    else cutout_kernel<float><<<grid, kThreads, 0, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
