// Kernel 3 — predicted-centre vote / group / NMS with bit-exact indices.
//
// Replaces nms_predicted_center (/root/reference/src/utils/utils.py:535-571)
// and its helpers canonical_to_global (:109-116), rphi_to_xy (:47-48).
//
// The reference sorts by confidence, builds a dense N x N float64 distance
// matrix and runs a Python loop.  The equivalent statement used here
// (oracle/nms.py::nms_sweep_spec, proven equal to the reference in the tests):
//   1. votes -> global (x, y) with NumPy's dtype promotion for the given input
//      dtypes; order = descending confidence (ties: higher index first);
//   2. adj[a][c] = dist(a, c) < min_dist in SORTED space, as an N x ceil(N/32)
//      bit matrix (embarrassingly parallel, L2 resident: 150 KB per JRDB scan);
//   3. ONE serial sweep: `if !suppressed[a]: keep a; suppressed |= adj[a]`, done
//      by a single warp, 32 candidates per step (the 32x32 diagonal block is
//      resolved with shuffles, kept rows are OR-ed into the remaining words by
//      all lanes in parallel);
//   4. instance id of point c = id of the LAST kept centre adjacent to c
//      (the reference's overwrite order, :565), found by scanning c's row of
//      `adj & keep` from the top.
//
// Three launches per call, each over all B scans; nothing returns to the host.
#include <math.h>

#include "pof_common.cuh"

namespace pof {
namespace {

constexpr int kMaxPoints = 4096;     // in-CTA bitonic sort capacity (32 KB of keys)
constexpr int kSortThreads = 1024;
constexpr int kAdjThreads = 256;
constexpr int kSweepThreads = 256;

struct NmsArgs {
    const void* scan;   // [B,N] float|double
    const void* phi;    // [N]   float|double
    const float* cls;   // [B,N]
    const float* reg;   // [B,N,2]
    int B, N, n_words;
    double min_dist;
    int* order;
    int* keep_idx;
    int* n_keep;
    int* instance_mask;
    double* det_xy;
    float* det_cls;
    double* xs;         // ws: [B,N] sorted-space x (exact widening when the compute type is float)
    double* ys;         // ws: [B,N]
    unsigned* adj;      // ws: [B,N,n_words]
};

// Total order on float32 confidences as unsigned keys (larger float -> larger key).
__device__ __forceinline__ unsigned orderable(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Elementary functions in the compute type.  For float they are evaluated in
// double and rounded once, i.e. correctly rounded up to double-rounding cases
// (NumPy's float32 kernels are within 1-2 ulp of that).
template <typename T> struct Fn;
template <> struct Fn<double> {
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    static __device__ __forceinline__ double cos_(double v) { return cos(v); }
    static __device__ __forceinline__ double sin_(double v) { return sin(v); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt_(double a) { return __dsqrt_rn(a); }
};
template <> struct Fn<float> {
    static __device__ __forceinline__ float atan2_(float y, float x) { return (float)atan2((double)y, (double)x); }
    static __device__ __forceinline__ float cos_(float v) { return (float)cos((double)v); }
    static __device__ __forceinline__ float sin_(float v) { return (float)sin((double)v); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }
};

template <bool A, typename X, typename Y> struct Select { typedef X type; };
template <typename X, typename Y> struct Select<false, X, Y> { typedef Y type; };

// ---- 1. votes -> xy, sort by descending confidence -------------------------------------------
template <bool SCAN64, bool PHI64>
__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const NmsArgs a) {
    typedef typename Select<SCAN64, double, float>::type T1;                 // dtype(scan)
    typedef typename Select<SCAN64 || PHI64, double, float>::type T2;        // promote(scan, phi)
    typedef typename Select<PHI64, double, float>::type PhiT;
    extern __shared__ unsigned long long keys[];   // n_pad entries
    const int b = blockIdx.x;
    const int N = a.N;
    int n_pad = 1;
    while (n_pad < N) n_pad <<= 1;

    const float* cls = a.cls + (size_t)b * N;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
        keys[i] = (i < N) ? (((unsigned long long)orderable(cls[i]) << 32) | (unsigned)i) : 0ull;
    __syncthreads();

    // bitonic sort, descending (padding keys are 0 = smaller than any real key, because
    // orderable() of any float sets at least one of the top 32 bits... except -NaN; fine)
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long ki = keys[i], kl = keys[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? (ki < kl) : (ki > kl)) { keys[i] = kl; keys[l] = ki; }
                }
            }
            __syncthreads();
        }
    }

    const T1* scan = reinterpret_cast<const T1*>(a.scan) + (size_t)b * N;
    const PhiT* phi = reinterpret_cast<const PhiT*>(a.phi);
    const float* reg = a.reg + (size_t)b * N * 2;
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const int i = (int)(keys[r] & 0xffffffffu);
        const float dx = reg[2 * i], dy = reg[2 * i + 1];
        const T1 fwd = Fn<T1>::add(scan[i], (T1)dy);                         // utils.py:110
        const T1 bearing = Fn<T1>::atan2_((T1)dx, fwd);                      // :111
        const T2 phi_v = Fn<T2>::add((T2)bearing, (T2)phi[i]);               // :114
        const T1 r_v = Fn<T1>::div(fwd, Fn<T1>::cos_(bearing));              // :115
        const T2 x = Fn<T2>::mul((T2)r_v, Fn<T2>::cos_(phi_v));              // :48
        const T2 y = Fn<T2>::mul((T2)r_v, Fn<T2>::sin_(phi_v));
        a.order[(size_t)b * N + r] = i;
        a.xs[(size_t)b * N + r] = (double)x;
        a.ys[(size_t)b * N + r] = (double)y;
    }
}

// ---- 2. adjacency bit matrix in sorted space --------------------------------------------------
// D[a, c] = sqrt((x_a - x_c)^2 + (y_a - y_c)^2) < min_dist in the promoted dtype (:550-552, :562), every operation
// rounded separately as NumPy does.  The square root is monotone, so the comparison is made on the sum of squares
// against `s_max`, the largest value whose ROUNDED root is still below the threshold (found once per CTA by
// stepping ulps around min_dist^2): the same boolean for every input, without N^2 square roots.
// A CTA stages the scan's centres in shared memory in the compute dtype (the float64 -> float32 conversion used to
// be done N^2 times from global memory); a thread owns one 32-column word of one row and visits its columns in an
// order rotated by its lane index, which makes the shared-memory reads bank-conflict free.
template <typename T> __device__ __forceinline__ T next_after(T v, int dir);
template <> __device__ __forceinline__ float next_after<float>(float v, int dir) {           // v >= 0
    if (v == 0.f) return dir > 0 ? __int_as_float(1) : -__int_as_float(1);
    return __int_as_float(__float_as_int(v) + dir);
}
template <> __device__ __forceinline__ double next_after<double>(double v, int dir) {
    if (v == 0.0) return dir > 0 ? __longlong_as_double(1) : -__longlong_as_double(1);
    return __longlong_as_double(__double_as_longlong(v) + dir);
}
template <typename T2>
__device__ T2 largest_square_below(T2 thr) {
    if (!(thr > (T2)0)) return (T2)-1;                       // nothing is closer than a non-positive distance
    T2 s = Fn<T2>::mul(thr, thr);
    if (!(s < (T2)3.0e38)) return s;                         // infinite threshold: everything finite qualifies
    for (int i = 0; i < 64 && Fn<T2>::sqrt_(s) < thr; ++i) s = next_after<T2>(s, +1);
    for (int i = 0; i < 128 && !(Fn<T2>::sqrt_(s) < thr); ++i) s = next_after<T2>(s, -1);
    return s;                                                 // sqrt_rn(s) < thr and sqrt_rn(next(s)) >= thr
}

template <typename T2>
__global__ void __launch_bounds__(kAdjThreads) nms_adjacency_kernel(const NmsArgs a) {
    extern __shared__ __align__(16) unsigned char adj_smem[];
    __shared__ T2 s_max_sh;
    const int b = blockIdx.y;
    const int N = a.N, nw = a.n_words;
    T2* sx = reinterpret_cast<T2*>(adj_smem);
    T2* sy = sx + N;
    const double* xs = a.xs + (size_t)b * N;
    const double* ys = a.ys + (size_t)b * N;
    unsigned* adj = a.adj + (size_t)b * N * nw;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { sx[i] = (T2)xs[i]; sy[i] = (T2)ys[i]; }
    if (threadIdx.x == 0) s_max_sh = largest_square_below<T2>((T2)a.min_dist);
    __syncthreads();
    const T2 s_max = s_max_sh;
    const int lane = threadIdx.x & 31;
    const long long total = (long long)N * nw;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(t / nw);
        const int w = (int)(t - (long long)row * nw);
        const T2 xa = sx[row], ya = sy[row];
        const int c0 = w << 5;
        unsigned bits = 0;
        if (c0 + 32 <= N) {
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const int k = (i + lane) & 31;
                const T2 ddx = Fn<T2>::sub(xa, sx[c0 + k]);                      // :550-552
                const T2 ddy = Fn<T2>::sub(ya, sy[c0 + k]);
                const T2 q = Fn<T2>::add(Fn<T2>::mul(ddx, ddx), Fn<T2>::mul(ddy, ddy));
                bits |= (q <= s_max ? 1u : 0u) << k;                             // :562
            }
        } else {
            for (int i = 0; i < 32; ++i) {
                const int k = (i + lane) & 31;
                if (c0 + k < N) {
                    const T2 ddx = Fn<T2>::sub(xa, sx[c0 + k]);
                    const T2 ddy = Fn<T2>::sub(ya, sy[c0 + k]);
                    const T2 q = Fn<T2>::add(Fn<T2>::mul(ddx, ddx), Fn<T2>::mul(ddy, ddy));
                    bits |= (q <= s_max ? 1u : 0u) << k;
                }
            }
        }
        adj[t] = bits;
    }
}

// ---- 3 + 4. serial sweep, ids, outputs ---------------------------------------------------------
__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const NmsArgs a) {
    __shared__ unsigned sup[kMaxPoints / 32];     // suppressed so far
    __shared__ unsigned kept[kMaxPoints / 32];    // kept centres
    __shared__ int kept_before[kMaxPoints / 32 + 1];
    const int b = blockIdx.x;
    const int N = a.N, nw = a.n_words;
    const unsigned* adj = a.adj + (size_t)b * N * nw;
    const int lane = threadIdx.x & 31;

    for (int w = threadIdx.x; w < nw; w += blockDim.x) { sup[w] = 0; kept[w] = 0; }
    __syncthreads();

    if (threadIdx.x < 32) {
        for (int blk = 0; blk < nw; ++blk) {
            const int cand = (blk << 5) + lane;
            const unsigned diag = (cand < N) ? adj[(size_t)cand * nw + blk] : 0u;
            unsigned dead = sup[blk];
            if (((blk + 1) << 5) > N) dead |= ~0u << (N - (blk << 5));       // padding candidates
            unsigned keep_bits = 0;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const unsigned row_l = __shfl_sync(0xffffffffu, diag, l);
                if (!((dead >> l) & 1u)) { keep_bits |= 1u << l; dead |= row_l; }
            }
            if (lane == 0) kept[blk] = keep_bits;
            // OR the kept rows into the words still to be visited
            for (int w = blk + 1 + lane; w < nw; w += 32) {
                unsigned acc = sup[w];
                unsigned todo = keep_bits;
                while (todo) {
                    const int l = __ffs(todo) - 1;
                    todo &= todo - 1;
                    acc |= adj[(size_t)((blk << 5) + l) * nw + w];
                }
                sup[w] = acc;
            }
            __syncwarp();
        }
        if (lane == 0) {
            int run = 0;
            for (int w = 0; w < nw; ++w) { kept_before[w] = run; run += __popc(kept[w]); }
            kept_before[nw] = run;
            a.n_keep[b] = run;
        }
    }
    __syncthreads();

    const int* order = a.order + (size_t)b * N;
    const float* cls = a.cls + (size_t)b * N;
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const int w_r = r >> 5, bit_r = r & 31;
        const int pt = order[r];
        if ((kept[w_r] >> bit_r) & 1u) {
            const int rank = kept_before[w_r] + __popc(kept[w_r] & ((1u << bit_r) - 1u));
            a.keep_idx[(size_t)b * N + rank] = pt;                                    // :568-569
            a.det_xy[((size_t)b * N + rank) * 2] = a.xs[(size_t)b * N + r];
            a.det_xy[((size_t)b * N + rank) * 2 + 1] = a.ys[(size_t)b * N + r];
            a.det_cls[(size_t)b * N + rank] = cls[pt];
        }
        int id = 0;
        for (int w = nw - 1; w >= 0; --w) {
            const unsigned m = adj[(size_t)r * nw + w] & kept[w];
            if (m) {
                const int top = 31 - __clz(m);
                id = kept_before[w] + __popc(kept[w] & ((2u << top) - 1u));           // 1-based
                break;
            }
        }
        a.instance_mask[(size_t)b * N + pt] = id;                                     // :565
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace pof

extern "C" {

size_t pof_nms_ws_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    const size_t nw = (size_t)(N + 31) / 32;
    return pof::align_up((size_t)B * N * sizeof(double), 256) * 2 + pof::align_up((size_t)B * N * nw * sizeof(unsigned), 256);
}

int pof_nms_centers(const void* scan, int scan_is_f64, const void* phi, int phi_is_f64, const float* cls,
                    const float* reg, int B, int N, double min_dist, int* order, int* keep_idx, int* n_keep,
                    int* instance_mask, double* det_xy, float* det_cls, void* ws, size_t ws_bytes, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(scan && phi && cls && reg && order && keep_idx && n_keep && instance_mask && det_xy && det_cls,
                POF_ERR_NULL_POINTER, "pof_nms_centers: null pointer argument");
    POF_REQUIRE(B >= 0 && N >= 1 && N <= kMaxPoints, POF_ERR_BAD_SHAPE, "pof_nms_centers: need 1 <= N <= %d (got %d)",
                kMaxPoints, N);
    POF_REQUIRE(min_dist == min_dist, POF_ERR_BAD_PARAM, "pof_nms_centers: min_dist is NaN");
    if (B == 0) return POF_OK;
    POF_REQUIRE(ws && ws_bytes >= pof_nms_ws_bytes(B, N), POF_ERR_WORKSPACE, "pof_nms_centers: workspace too small (%zu < %zu)",
                ws_bytes, pof_nms_ws_bytes(B, N));

    NmsArgs a;
    a.scan = scan; a.phi = phi; a.cls = cls; a.reg = reg;
    a.B = B; a.N = N; a.n_words = (N + 31) / 32;
    a.min_dist = min_dist;
    a.order = order; a.keep_idx = keep_idx; a.n_keep = n_keep; a.instance_mask = instance_mask;
    a.det_xy = det_xy; a.det_cls = det_cls;
    char* p = reinterpret_cast<char*>(ws);
    const size_t xy_bytes = align_up((size_t)B * N * sizeof(double), 256);
    a.xs = reinterpret_cast<double*>(p);
    a.ys = reinterpret_cast<double*>(p + xy_bytes);
    a.adj = reinterpret_cast<unsigned*>(p + 2 * xy_bytes);

    int n_pad = 1;
    while (n_pad < N) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(unsigned long long);
    const int sort_threads = n_pad / 2 < kSortThreads ? (n_pad / 2 < 32 ? 32 : n_pad / 2) : kSortThreads;
    if (scan_is_f64 && phi_is_f64) nms_sort_kernel<true, true><<<B, sort_threads, smem, stream>>>(a);
    else if (scan_is_f64) nms_sort_kernel<true, false><<<B, sort_threads, smem, stream>>>(a);
    else if (phi_is_f64) nms_sort_kernel<false, true><<<B, sort_threads, smem, stream>>>(a);
    else nms_sort_kernel<false, false><<<B, sort_threads, smem, stream>>>(a);
    POF_CUDA(cudaGetLastError());

    const long long words = (long long)N * a.n_words;
    // a few CTAs per scan (each stages the scan's centres once), enough scans x CTAs to fill the SMs
    unsigned gx = (unsigned)((words + kAdjThreads * 16 - 1) / (kAdjThreads * 16));
    dim3 grid_adj(gx > 0 ? gx : 1, (unsigned)B);
    if ((scan_is_f64 || phi_is_f64) && (size_t)N * 2 * sizeof(double) > 48 * 1024)
        POF_CUDA(cudaFuncSetAttribute(nms_adjacency_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kMaxPoints * 2 * (int)sizeof(double)));
    if (scan_is_f64 || phi_is_f64) nms_adjacency_kernel<double><<<grid_adj, kAdjThreads, (size_t)N * 2 * sizeof(double), stream>>>(a);
    else nms_adjacency_kernel<float><<<grid_adj, kAdjThreads, (size_t)N * 2 * sizeof(float), stream>>>(a);
    POF_CUDA(cudaGetLastError());

    nms_sweep_kernel<<<B, kSweepThreads, 0, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
