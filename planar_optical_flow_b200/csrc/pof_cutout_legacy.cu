// Legacy preprocessing of the reference: used when a config has no `area_mode`
// key (src/utils/dataset_dr_spaam.py:440-443; config/config_cluster.yaml, config_obj_det.yaml) or
// selects the `fc2d` network.
//
//   pof_cutout_original_fwd   scans_to_cutout_original, /root/reference/src/utils/utils.py:423-489:
//       integer beam window [round(i - ha/dphi), round(i + ha/dphi)] (out-of-scan beams read the padding
//       value), resampled to P points with OpenCV's cv2.resize — INTER_AREA when the window is longer than
//       P, INTER_LINEAR otherwise — then depth clip / centring in float32.  The resize arithmetic restates
//       OpenCV's imgproc/resize.cpp for a one-column float image (resizeGeneric_ + VResizeLinear,
//       ResizeAreaFast_, computeResizeAreaTab + ResizeArea_): same tables (double), same float32
//       products and sums in the same order, no contraction.  One warp per (scan, point) row, lanes
//       along the P output samples.
//   pof_polar_grid_fwd        scans_to_polar_grid, utils.py:492-531 (truncated signed distance along
//       the range axis, the measured range written into its own bin).  One thread per output element.
#include <float.h>
#include <math.h>

#include "pof_common.cuh"

namespace pof {
namespace {

struct LegacyArgs {
    const float* scans;   // [B, S, N]
    float* out;           // [B, N, S, P]
    int B, S, N, P;
    double incre_d;       // beam pitch (double, or the float32 value widened when incre_is_f32)
    int incre_is_f32;
    float half_width_f;   // (float)(0.5 * window_width)
    double half_width_d;  // 0.5 * window_width
    float depth_f, pad_f;
    int fixed, centered;
};

// round half to even, like Python's round() on a float
__device__ __forceinline__ long long py_round(double v) { return __double2ll_rn(v); }

__device__ __forceinline__ float fetch(const float* scan, int N, int start, int k, float pad) {
    const int g = min(max(start + k, -1), N);                         // utils.py:455
    return (g < 0 || g >= N) ? pad : __ldg(scan + g);                 // index -1 and N both hit the padding column
}

__global__ void __launch_bounds__(256) cutout_original_kernel(const LegacyArgs a) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= (long long)a.B * a.S * a.N) return;
    const int i = (int)(row % a.N);
    const int bs = (int)(row / a.N), b = bs / a.S, s = bs - b * a.S;
    const float* scan = a.scans + (size_t)bs * a.N;
    const float pt_r = a.fixed ? __ldg(scan + i) : __ldg(a.scans + ((size_t)b * a.S + (a.S - 1)) * a.N + i);    // :448

    // :450  float(np.arctan(0.5 * window_width / max(pt_r, 0.01))) under NumPy's scalar promotion
    double ha;
    if (pt_r < 0.01f) ha = atan(a.half_width_d / 0.01);                // python floats all the way
    else ha = (double)(float)atan((double)__fdiv_rn(a.half_width_f, pt_r));   // float32 ratio, float32 arctangent
    long long start, end;                                             // :452-453
    if (a.incre_is_f32) {
        const float q = __fdiv_rn((float)ha, (float)a.incre_d);
        start = py_round((double)__fsub_rn((float)i, q));
        end = py_round((double)__fadd_rn((float)i, q));
    } else {
        const double q = __ddiv_rn(ha, a.incre_d);
        start = py_round(__dsub_rn((double)i, q));
        end = py_round(__dadd_rn((double)i, q));
    }
    const int n = (int)(end - start + 1);                             // window length, >= 1
    const int st = (int)start;
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)a.P, (double)n));
    const bool area = a.P < n;                                        // :464-468
    const double rscale = rint(scale);
    const bool area_fast = area && fabs(scale - rscale) < DBL_EPSILON;
    const float lo = __fsub_rn(pt_r, a.depth_f), hi = __fadd_rn(pt_r, a.depth_f);     // :474-476
    float* dst = a.out + (((size_t)b * a.N + i) * a.S + s) * a.P;

    for (int dy = lane; dy < a.P; dy += 32) {
        float v;
        if (area_fast) {                                              // ResizeAreaFast_: plain sum, times 1/area
            const int is = (int)rscale;
            float acc = 0.f;
            for (int k = 0; k < is; ++k) acc = __fadd_rn(acc, fetch(scan, a.N, st, dy * is + k, a.pad_f));
            v = __fmul_rn(acc, __fdiv_rn(1.f, (float)is));
        } else if (area) {                                            // computeResizeAreaTab + ResizeArea_
            const double f1 = dy * scale, f2 = f1 + scale;
            const double cell = fmin(scale, (double)n - f1);
            int s1 = (int)ceil(f1), s2 = (int)floor(f2);
            s2 = min(s2, n - 1);
            s1 = min(s1, s2);
            float acc = 0.f;
            bool first = true;
            if ((double)s1 - f1 > 1e-3) {
                acc = __fmul_rn((float)(((double)s1 - f1) / cell), fetch(scan, a.N, st, s1 - 1, a.pad_f));
                first = false;
            }
            const float full = (float)(1.0 / cell);
            for (int sx = s1; sx < s2; ++sx) {
                const float t = __fmul_rn(full, fetch(scan, a.N, st, sx, a.pad_f));
                acc = first ? t : __fadd_rn(acc, t);
                first = false;
            }
            if (f2 - (double)s2 > 1e-3) {
                const float t = __fmul_rn((float)(fmin(fmin(f2 - (double)s2, 1.0), cell) / cell), fetch(scan, a.N, st, s2, a.pad_f));
                acc = first ? t : __fadd_rn(acc, t);
            }
            v = acc;
        } else {                                                      // resizeGeneric_ linear, border replicated
            float fy = (float)(((double)dy + 0.5) * scale - 0.5);
            const int sy = (int)floorf(fy);
            fy = __fsub_rn(fy, (float)sy);
            const float v0 = fetch(scan, a.N, st, min(max(sy, 0), n - 1), a.pad_f);
            const float v1 = fetch(scan, a.N, st, min(max(sy + 1, 0), n - 1), a.pad_f);
            v = __fadd_rn(__fmul_rn(v0, __fsub_rn(1.f, fy)), __fmul_rn(v1, fy));
        }
        v = fminf(fmaxf(v, lo), hi);
        if (a.centered) v = __fdiv_rn(__fsub_rn(v, pt_r), a.depth_f);  // :485-486
        dst[dy] = v;
    }
}

struct PolarArgs {
    const float* scans;   // [S, N]
    float* out;           // [S, R, N]
    int S, N, R;
    float min_f, max_f, bin_f, clip_f, mid_f, mag_f;
    int use_tsdf, normalize;
};

__global__ void __launch_bounds__(256) polar_grid_kernel(const PolarArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)a.S * a.R * a.N) return;
    const int i = (int)(t % a.N);
    const int r = (int)((t / a.N) % a.R);
    const int s = (int)(t / ((long long)a.N * a.R));
    const float val = fminf(fmaxf(__ldg(a.scans + (size_t)s * a.N + i), a.min_f), a.max_f);          // :505
    const int ind = (int)__fdiv_rn(__fsub_rn(val, a.min_f), a.bin_f);                                 // :506
    float v;
    if (r == ind) {
        v = a.normalize ? __fmul_rn(__fdiv_rn(__fsub_rn(val, a.mid_f), a.mag_f), 2.f) : val;          // :523,527
    } else {
        v = a.use_tsdf ? fminf(fmaxf(__fmul_rn((float)(r - ind), a.bin_f), -a.clip_f), a.clip_f) : 0.f;   // :513-520
        if (a.normalize) v = __fmul_rn(__fdiv_rn(v, a.mag_f), 2.f);                                   // :525
    }
    a.out[t] = v;
}

}  // namespace
}  // namespace pof

extern "C" {

int pof_cutout_original_fwd(const float* scans, int B, int S, int N, double angle_incre, int angle_incre_is_f32, int P,
                            double window_width, double window_depth, double padding_val, int fixed, int centered, float* out,
                            void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B == 0) return POF_OK;
    POF_REQUIRE(scans && out, POF_ERR_NULL_POINTER, "pof_cutout_original_fwd: null pointer");
    POF_REQUIRE(B > 0 && S >= 1 && N >= 1 && P >= 1, POF_ERR_BAD_SHAPE, "pof_cutout_original_fwd: bad shape B=%d S=%d N=%d P=%d", B, S, N, P);
    POF_REQUIRE(angle_incre > 0.0 && window_depth > 0.0, POF_ERR_BAD_PARAM, "pof_cutout_original_fwd: angle_incre and window_depth must be positive");
    LegacyArgs a;
    a.scans = scans; a.out = out; a.B = B; a.S = S; a.N = N; a.P = P;
    a.incre_d = angle_incre; a.incre_is_f32 = angle_incre_is_f32;
    a.half_width_d = 0.5 * window_width; a.half_width_f = (float)a.half_width_d;
    a.depth_f = (float)window_depth; a.pad_f = (float)padding_val;
    a.fixed = fixed; a.centered = centered;
    const long long rows = (long long)B * S * N;
    cutout_original_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

int pof_polar_grid_fwd(const float* scans, int S, int N, double min_range, double max_range, double range_bin_size, double tsdf_clip,
                       int normalize, float* out, void* stream_) {
    using namespace pof;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (S == 0) return POF_OK;
    POF_REQUIRE(scans && out, POF_ERR_NULL_POINTER, "pof_polar_grid_fwd: null pointer");
    POF_REQUIRE(S > 0 && N >= 1 && range_bin_size > 0.0 && max_range > min_range, POF_ERR_BAD_PARAM, "pof_polar_grid_fwd: bad arguments");
    PolarArgs a;
    a.scans = scans; a.out = out; a.S = S; a.N = N;
    a.R = (int)((max_range - min_range) / range_bin_size) + 1;        // :501
    a.min_f = (float)min_range; a.max_f = (float)max_range; a.bin_f = (float)range_bin_size; a.clip_f = (float)tsdf_clip;
    a.mag_f = (float)(max_range - min_range); a.mid_f = (float)(0.5 * (max_range - min_range));
    a.use_tsdf = tsdf_clip > 0.0; a.normalize = normalize;
    const long long total = (long long)S * a.R * N;
    polar_grid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

}  // extern "C"
