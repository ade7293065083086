// Shared helpers for libpof.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pof.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpof is written for sm_100a (B200); build with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace pof {

// Thread-local message returned by pof_last_error().
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define POF_REQUIRE(cond, code, ...)          \
    do {                                      \
        if (!(cond)) {                        \
            ::pof::set_error(__VA_ARGS__);    \
            return (code);                    \
        }                                     \
    } while (0)

#define POF_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return ::pof::cuda_fail(e__, #call); \
    } while (0)

// B200: 148 SMs.  Queried once per device; only used to size persistent grids.
int sm_count();

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Streaming (read-once / write-once) 128-bit accesses: keep them out of L1 so the
// small re-used tables (scan rows, embeddings, weights) stay resident there.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace pof
