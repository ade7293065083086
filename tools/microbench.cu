// Instruction-throughput microbenchmarks that decide the cutout kernel's arithmetic plan.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o /tmp/microbench tools/microbench.cu   (build outside the tree: the binary is not part of the product)
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 4

template <int OP>
__global__ void k(double* out, double seed, float fseed, int iseed) {
    double a[CHAINS];
    float f[CHAINS];
    int n[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { a[c] = seed + threadIdx.x * 1e-3 + c; f[c] = fseed + threadIdx.x + c; n[c] = iseed + threadIdx.x + c; }
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) a[c] = fma(a[c], 1.0000001, 0.5);                       // DFMA
            if (OP == 1) a[c] = __dadd_rn(a[c], 0.5);                            // DADD
            if (OP == 2) { a[c] = (double)f[c]; f[c] += (float)i; }              // F2F.F64.F32 (+FADD)
            if (OP == 3) { f[c] = (float)a[c]; a[c] = __dadd_rn(a[c], 0.5); }    // F2F.F32.F64 (+DADD)
            if (OP == 4) { a[c] = (double)n[c]; n[c] += i; }                     // I2F.F64 (+IADD)
            if (OP == 5) { n[c] = __double2int_rd(a[c]); a[c] = __dadd_rn(a[c], 0.5); }   // F2I.F64 floor (+DADD)
            if (OP == 6) a[c] = floor(a[c]) + 0.5;                               // FRND.F64 (+DADD)
            if (OP == 7) a[c] = __ddiv_rn(a[c], 1.0000001);                      // full double division
            if (OP == 8) f[c] = fmaf(f[c], 1.0000001f, 0.5f);                    // FFMA
            if (OP == 9) { a[c] = __dadd_rd(a[c], 4503599627370496.0); a[c] = __dadd_rn(a[c], -4503599627370495.5); }  // magic floor (2 DADD)
            if (OP == 10) a[c] = fmin(fmax(a[c], 0.25), 1e300) + 0.5;            // double clamp (+DADD)
            if (OP == 11) { n[c] = __float2int_rd(f[c]); f[c] += 0.5f; }         // F2I.F32 (+FADD)
            if (OP == 12) { f[c] = (float)n[c]; n[c] += i; }                     // I2F.F32 (+IADD)
            if (OP == 13) a[c] = __dmul_rn(a[c], 1.0000001);                     // DMUL
            if (OP == 14) f[c] = atanf(f[c]) + 1.0f;                             // atanf
            if (OP == 15) a[c] = atan(a[c]) + 1.0;                               // atan double
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += a[c] + f[c] + n[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double* out, int extra_ops) {
    const int blocks = 148 * 8, threads = 256;
    k<OP><<<blocks, threads>>>(out, 1.0, 1.0f, 1);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.0, 1.0f, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * ITER * CHAINS;
    // lanes per clock per SM, assuming ~1.9 GHz
    printf("%-34s %8.3f ms  %8.1f Gop/s  ~%6.1f lane-ops/clk/SM (@1.9GHz)  [%d helper op(s) per iteration included]\n", name, ms,
           ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.9e9, extra_ops);
}

int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    run<8>("FFMA", out, 0);
    run<0>("DFMA", out, 0);
    run<1>("DADD", out, 0);
    run<13>("DMUL", out, 0);
    run<2>("F2F.F64.F32 + FADD", out, 1);
    run<3>("F2F.F32.F64 + DADD", out, 1);
    run<4>("I2F.F64.S32 + IADD", out, 1);
    run<5>("F2I.S32.F64.FLOOR + DADD", out, 1);
    run<6>("floor(double) + DADD", out, 1);
    run<9>("magic floor (DADD.RM + DADD)", out, 0);
    run<7>("ddiv_rn", out, 0);
    run<10>("double clamp + DADD", out, 1);
    run<11>("F2I.S32.F32 + FADD", out, 1);
    run<12>("I2F.F32.S32 + IADD", out, 1);
    run<14>("atanf + FADD", out, 1);
    run<15>("atan(double) + DADD", out, 1);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
