// fp32-accurate 1-D convolution / whole-row GEMM on the sm_100a tensor cores.
//
// The DROW / SpatialDROW backbone (/root/reference/src/depracted/model/dr_spaam.py:8-12, 49-59, 87-114)
// is eleven Conv1d(k = 3, p = 1) + BatchNorm + LeakyReLU layers plus the gate's Conv1d(k = L) embedding
// (:130-133).  With BatchNorm folded and channels-last activations every one of them is
//
//     out[m, l, n] = sum_{t < taps} sum_{c < Cin}  A[m, l + t - pad, c] * W[t][n, c]        (zero outside 0 <= l' < LA)
//
// i.e. `taps` accumulated GEMMs over row-shifted views of the same matrix.  This kernel runs them on
// tcgen05 with the operand split x = hi + lo (hi exactly TF32, lo = the rounded remainder; "3xTF32"):
//
//     x*w ~= lo*w_hi + hi*w_lo + hi*w_hi                      (dropped lo*w_lo: 2^-22 relative)
//
// Why not cuDNN with the same split (engine precision "tf32x3")?  The tensor core accumulates with
// TRUNCATION: every 8-deep k-step loses ~2^-24 of the running sum, always towards zero, and over the
// 4608-deep reductions of this network that is a coherent 1e-4 bias (profiles/r1_precision_modes.txt).
// Here a reduction is cut into CHAINS of 64 channels of one tap (four 16-channel k-blocks): inside a
// k-block the two small correction products are issued before the main product, and each finished chain
// is read back from tensor memory and added to a register accumulator with a rounded fp32 add.  Measured
// against fp64 (profiles/r1_conv_tc_layers_v3.txt): 5-7e-7 per layer with 64-channel chains, 3e-7 with 32,
// against 1-2e-6 for cuDNN's fp32 SIMT kernels on the same layers.
//
// Structure (one persistent CTA per SM, 384 threads; an SM pair works on 256 rows x BN output channels with
// cta_group::2, each CTA holding its own 128 rows and half of the weight tile):
//   warp 0     TMA producer: per k-block (one tap, one 128-byte row of input channels: 64 binary16 / 32 TF32) four
//              tensor-map loads into a shared-memory ring - A_hi / A_lo as 3-D boxes (channels, Lout rows, mt cutouts)
//              whose row coordinate is shifted by the tap, so the convolution's zero padding is the TMA's
//              out-of-bounds fill; W_hi / W_lo as 2-D boxes (or, for the narrow layers, resident for the whole
//              launch).  128-byte swizzle, completion on an mbarrier (complete_tx).  The ring takes as many stages
//              as fit in 205 KB (3 at BN = 256, up to 12 with resident weights).
//   warp 1     MMA issuer (the pair's leader only): per k-block 3 products x (128 / 32) k-steps of tcgen05.mma
//              (kind::f16, K = 16, or kind::tf32, K = 8) into one of two tensor-memory accumulators; tcgen05.commit
//              releases the ring slot and, at the end of a chain, hands the accumulator to the epilogue warps.
//   warp 2     tensor-memory allocation (512 columns) and release.
//   warps 4-11 promotion + epilogue: tcgen05.ld the finished chain, add into registers, hand the TMEM
//              buffer back; after the last chain: scale + bias, LeakyReLU, max-pool over row pairs (shuffle),
//              optional hi/lo split for the next layer, transposition through shared memory, 128-bit stores.
// What bounds it (profiles/r1_conv_tc_f16_ncu_256to512.txt): the wide layers run the tensor
// pipe at ~60 % next to 104 of 128 B/clk of shared-memory traffic; the narrow layers stream their split
// activations from and to HBM.
// Every mbarrier wait is bounded: on a timeout the role records an error code and leaves, so a bug
// surfaces as a status, never as a hung GPU.
//
// F16 = true is the same pipeline on kind::f16: x = hi + lo with BOTH parts binary16 (11-bit significands like
// TF32, so the three products carry the same 22 bits), at twice the tensor-core rate and half the operand
// bytes - a 64-byte swizzled row now holds 32 channels instead of 16, which halves the shared-memory traffic
// that bounds the narrow layers and the HBM traffic of the split activations.  binary16's range is handled by
// scaling, not by a second accumulator: the weights of a layer are multiplied by a power of two that brings
// their largest magnitude to 2^13 (hi and lo are then both normal numbers) and the epilogue multiplies the
// sum by the inverse power (exact); activations are stored unscaled - below 2^-3 their lo part is subnormal,
// i.e. an ABSOLUTE error of 2^-25 on an O(1) activation - and an activation beyond 65504 raises status 16.
#include <cuda.h>
#include <cuda_fp16.h>

#include "pof_common.cuh"

namespace pof {
namespace {

constexpr int kThreadsBase = 128;               // warps 0-3: TMA producer, MMA issuer, TMEM allocator, (idle)
constexpr int kTileM = 128;
#ifndef POF_CONV_L2_PROMO
#define POF_CONV_L2_PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
#ifndef POF_CONV_ROW_BYTES
#define POF_CONV_ROW_BYTES 128
#endif
constexpr int kRowBytes = POF_CONV_ROW_BYTES;  // one swizzled operand row (= the TMA box's inner extent): 64 or 128 bytes
static_assert(kRowBytes == 64 || kRowBytes == 128, "operand rows are 64 or 128 bytes");
template <bool F16> struct KBlock { static constexpr int value = kRowBytes / (F16 ? 2 : 4); };   // channels per k-block
constexpr int kATile = kTileM * kRowBytes;     // bytes of one A operand tile (hi or lo)
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 12;                 // barrier slots: full[12] | empty[12] | tfull[2] | tempty[2] | wfull | tmem pointer
constexpr long long kWaitLimit = 2000000000LL; // ~1 s of SM clocks

// CG = 1: one CTA computes a 128 x BN tile.  CG = 2: an SM pair (cta_group::2) computes 256 x BN, each CTA
// holding its 128 rows of A and HALF of the weight tile: the MMA then reads 64 B/clk of operands per SM
// instead of 96, which is what lets it run at full rate next to the TMA fill (shared memory moves 128 B/clk).
constexpr int kRingBytes = 205 * 1024;                 // operand ring (227 KB per CTA minus staging, barriers and alignment slack)
constexpr int kStagePitch = 20;                        // words per staged output row: 16 + 4 (conflict-free 16-byte accesses)
constexpr int kStagingWarp = 32 * kStagePitch * 4;      // one 32 x 16-word transposition buffer per epilogue warp

template <int BN, int CG>
struct Cfg {
    static constexpr int kBTile = (BN / CG) * kRowBytes;       // this CTA's share of one W operand tile (hi or lo)
    static constexpr int kStage = 2 * kATile + 2 * kBTile;     // streamed weights: a stage holds A hi/lo and W hi/lo of one k-block
    static constexpr int kStageA = 2 * kATile;                 // resident weights: a stage holds A hi/lo only
    // Epilogue warps: 8 (two column halves per TMEM lane quarter).  A 16-warp form (four column quarters, 96 registers
    // per thread) exists for the 128-column tile behind POF_CONV_EPI16: it was expected to help the narrow layers'
    // dependent epilogue code and measured 2-5 % slower.
#ifndef POF_CONV_EPI16
#define POF_CONV_EPI16 0     // measured on one GPU (A/B): 64->128 0.85 ms with 16 epilogue warps, 0.81 ms with 8
#endif
    static constexpr int kEpiWarps = (BN == 128 && POF_CONV_EPI16) ? 16 : 8;      // (BN = 64 would leave 16 columns per thread: keeps 8)
    static constexpr int kThreads = kThreadsBase + 32 * kEpiWarps;
    static constexpr int kSlices = kEpiWarps / 4;              // column slices per TMEM lane quarter
    static constexpr int kAcc = BN / kSlices;                  // accumulators per epilogue thread
    static constexpr int kRing = kRingBytes - (kEpiWarps - 8) * kStagingWarp;
    static constexpr int kSmem = kRing + 1024 /* alignment slack */ + 256 /* barriers */ + kEpiWarps * kStagingWarp;
};

struct Params {
    long long Mcut;      // cutouts (outer dimension of A)
    long long tiles_m;
    int tiles_n;
    int Lout;            // output rows per cutout (= box height)
    int mt;              // cutouts per tile, mt * Lout <= 128
    int halo;            // 1: a tile's cutouts are loaded ONCE with their two halo rows, the three taps are row-shifted views (see the kernel)
    int split_h0;        // 0, or (SM pairs only) the pair's 256 rows hold 2 mt + 1 cutouts: the middle one is split, its first
                         // split_h0 = 128 - mt * Lout rows go to the leader CTA's tile, the other Lout - split_h0 to the peer's
    int Cin, Cout, taps, pad, pool;
    int chain;           // k-blocks accumulated in tensor memory before a promotion to registers
    int w_resident;      // the CTA's share of ALL weight k-blocks stays in shared memory for the whole launch
    int stages;          // ring depth (<= kMaxStages)
    float slope;
    float out_scale;     // multiplies the accumulated sum before the bias (F16: the inverse of the weights' power of two)
    const float* bias;   // [Cout] or null
    float* out_plain;    // [Mcut * Lout / pool, Cout] or null
    void* out_split;     // [Mcut * Lout / pool, 2 Cout] = [hi | lo] (fp32, or binary16 when F16) or null
    int* status;         // device int, 0 = ok
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait; false = timed out (status written).
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, int* status, int code) {
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (mbar_try(bar, parity)) return true;
        if (clock64() - t0 > kWaitLimit) {
            atomicCAS(status, 0, code);
            return false;
        }
    }
}

// Whole-warp form of the bounded wait: every lane polls, the votes make the outcome (and the time-out) WARP-UNIFORM, so
// the compiler keeps the loop counters, ring addresses and matrix descriptors that depend on it in uniform registers.
// (With a single-lane role - `lane == 0` around the whole loop - nothing is provably uniform and every tcgen05.mma /
// TMA instruction is wrapped in a R2UR waterfall loop: 353 SASS instructions per k-block for 12 MMAs, and that one
// thread's issue rate, not the tensor pipe, bounded every layer.)
__device__ __forceinline__ bool mbar_wait_warp(unsigned bar, unsigned parity, int* status, int code) {
    if (__all_sync(0xffffffffu, mbar_try(bar, parity))) return true;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (__all_sync(0xffffffffu, mbar_try(bar, parity))) return true;
        if (__any_sync(0xffffffffu, clock64() - t0 > kWaitLimit)) {
            if ((threadIdx.x & 31) == 0) atomicCAS(status, 0, code);
            return false;
        }
    }
}
// One lane of a converged warp (the same lane every time for the same mask).
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Warm the L2 with a box the producer will load a tile later (the first touch of the activations comes from HBM).
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// Shared-memory matrix descriptor, K-major, swizzled rows of kRowBytes (128 B; 8-row groups 1024 B apart).
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr) {
    return (unsigned long long)((addr >> 4) & 0x3fffu) | (1ull << 16) /* LBO (unused with swizzle) */ |
           ((unsigned long long)(8 * kRowBytes >> 4) << 32) /* SBO */ | (1ull << 46) /* descriptor version (sm_100) */ |
           ((kRowBytes == 128 ? 2ull : 4ull) << 61) /* SWIZZLE_128B / SWIZZLE_64B */;
}
// Instruction descriptor: D = fp32, A = B = TF32 (format 2) or binary16 (format 0), both K-major, M = m, N = bn.
__device__ __forceinline__ unsigned umma_idesc(int bn, int m, bool f16) {
    const unsigned fmt = f16 ? 0u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(bn >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// ---- cta_group::2 forms: the leader CTA (cluster rank 0) issues for the pair -----------------------
constexpr unsigned kPeerBitMask = 0xFEFFFFFFu;    // clears the CTA-rank bit of a shared::cluster address -> the leader's copy
__device__ __forceinline__ void tma_load_3d_pair(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(unsigned bar) {      // arrives on the same barrier in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(unsigned bar) {    // arrive on the leader CTA's copy of `bar`
    asm volatile(
        "{\n\t.reg .b32 r;\n\t"
        "mapa.shared::cluster.u32 r, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [r];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"          // same statement: the registers are valid when it returns
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld64(unsigned taddr, unsigned (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
          "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
          "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
          "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
          "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
    unsigned u = __float_as_uint(x);
    u += 0x0fffu + ((u >> 13) & 1u);
    return __uint_as_float(u & 0xffffe000u);
}
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

template <int BN, int CG, bool F16>
__global__ void __cluster_dims__(CG, 1, 1) __launch_bounds__((Cfg<BN, CG>::kThreads), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_p1, const Params p) {
    using C = Cfg<BN, CG>;
    constexpr int kKBlock = KBlock<F16>::value;
    extern __shared__ unsigned char smem_raw[];
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned bars = base + C::kRing;                     // full[12] | empty[12] | tfull[2] | tempty[2] | wfull | tmem ptr
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (kMaxStages + s); };
    auto tfull = [&](int b) { return bars + 8u * (2 * kMaxStages + b); };
    auto tempty = [&](int b) { return bars + 8u * (2 * kMaxStages + 2 + b); };
    const unsigned wfull = bars + 8u * (2 * kMaxStages + 4);
    const unsigned tmem_slot = bars + 8u * (2 * kMaxStages + 5);
    float* staging = reinterpret_cast<float*>(smem_raw + (bars + 256u - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = CG > 1 ? cluster_rank() : 0u;          // rank 0 of a pair leads: it issues the MMAs
    const int kb_per_tap = p.Cin / kKBlock;
    const int n_kb = p.taps * kb_per_tap;
    // work items: (group of CG consecutive row tiles) x (column tile); a CTA takes row tile CG * g + rank
    const long long n_tiles = ((p.tiles_m + CG - 1) / CG) * p.tiles_n;
    const long long first_tile = blockIdx.x / CG, tile_stride = gridDim.x / CG;
    // Split tiles (p.split_h0 > 0): with Lout = 28 a CTA's 128 rows hold four cutouts and 16 idle rows; instead the pair's 256
    // rows take NINE cutouts - the leader's tile = cutouts 0-3 + rows [0, h0) of cutout 4, the peer's = cutouts 5-8 + rows
    // [h0, Lout) of cutout 4, each CTA loading its part of the middle cutout with its own box (the row shift of a tap crosses
    // the cut into real rows, and leaves the cutout into the zero fill, exactly like the full boxes).  In both CTAs the
    // whole cutouts come first in shared memory (the second box then starts 1024-byte aligned), so the peer's tile rows are not
    // in output order: its epilogue stores the two segments separately.
    const bool split = CG == 2 && p.split_h0 > 0;
    const int rows_full = p.mt * p.Lout;                         // rows of whole cutouts in a CTA's tile
    const int rows_tile = !split ? rows_full : rank == 0 ? rows_full + p.split_h0 : rows_full + p.Lout - p.split_h0;
    const int per_pair = split ? 2 * p.mt + 1 : CG * p.mt;       // cutouts per work item
    // Halo tiles (p.halo; Cin = one k-block, k = 3, the weights resident): the layers with 56 rows per cutout are bound by the
    // TMA's L2 -> shared-memory row rate, and a tile is loaded three times, once per tap.  Here a cutout comes in ONCE with its
    // two padding rows - box rows -1 .. Lout, zero filled by the TMA like any row outside the cutout - and tap t multiplies the
    // view that starts t rows further down: the descriptor's start address moves by t * 128 B (the 128-byte swizzle is a
    // function of the absolute shared-memory address, so a row-shifted view of a tile the TMA wrote reads the right chunks).
    // Accumulator lane i is then output row i - 2 c of cutout c = i / (Lout + 2); the two lanes between cutouts are not outputs.
    // Same MMAs in the same order as the tap-by-tap form (tap 0, tap 1 | tap 2: the chains of a 64-channel layer), so the
    // same bits.
    const bool halo = p.halo != 0;
    const int halo_rows = p.Lout + 2;
    // shared-memory plan: [resident weights: n_kb x (W hi | W lo)] [ring: stages x stage_bytes]
    const int n_stages = p.stages;
    const unsigned stage_bytes = p.w_resident ? C::kStageA : C::kStage;
    const unsigned ring = base + (p.w_resident ? (unsigned)n_kb * 2u * C::kBTile : 0u);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(wfull, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(tfull(b), 1); mbar_init(tempty(b), C::kEpiWarps * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG > 1) cluster_sync();          // the peer's barriers and tensor memory exist before anything is sent to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0) {
            // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
            const unsigned rows_pair = split ? (unsigned)(per_pair * p.Lout) : CG * (unsigned)rows_tile;
            const unsigned tx = 2u * rows_pair * kRowBytes + CG * (p.w_resident ? 0u : 2u * (unsigned)C::kBTile);   // both CTAs' loads land on the leader's barrier
            int s = 0;
            unsigned ph = 0;
            bool ok = true;
            if (p.w_resident && first_tile < n_tiles && elect_one()) {          // every weight k-block of this CTA's columns, once
                const int n0 = (int)rank * (BN / CG);
                if (rank == 0) mbar_expect_tx(wfull, CG * (unsigned)n_kb * 2u * (unsigned)C::kBTile);
                for (int kb = 0; kb < n_kb; ++kb) {
                    const int tap = kb / kb_per_tap, c0 = (kb - tap * kb_per_tap) * kKBlock;
                    const unsigned dst = base + (unsigned)kb * 2u * C::kBTile;
                    if (CG == 1) {
                        tma_load_2d(dst, &map_w, c0, (tap * 2 + 0) * p.Cout + n0, wfull);
                        tma_load_2d(dst + C::kBTile, &map_w, c0, (tap * 2 + 1) * p.Cout + n0, wfull);
                    } else {
                        tma_load_2d_pair(dst, &map_w, c0, (tap * 2 + 0) * p.Cout + n0, wfull);
                        tma_load_2d_pair(dst + C::kBTile, &map_w, c0, (tap * 2 + 1) * p.Cout + n0, wfull);
                    }
                }
            }
            __syncwarp();
            if (halo) {                                        // one box per part and tile: (channels, Lout + 2 rows from row -1, mt cutouts)
                const unsigned tx_h = CG * 2u * (unsigned)(p.mt * halo_rows) * kRowBytes;
                for (long long tile = first_tile; tile < n_tiles && ok; tile += tile_stride) {
                    const int m0 = (int)(((tile / p.tiles_n) * CG + rank) * p.mt);
                    if (!(ok = mbar_wait_warp(empty(s), ph ^ 1u, p.status, 1))) break;
                    const unsigned dst = ring + (unsigned)s * stage_bytes;
                    if (elect_one()) {
                        if (CG == 1) {
                            mbar_expect_tx(full(s), tx_h);
                            tma_load_3d(dst, &map_p0, 0, -1, m0, full(s));
                            tma_load_3d(dst + kATile, &map_p0, p.Cin, -1, m0, full(s));
                        } else {
                            if (rank == 0) mbar_expect_tx(full(s), tx_h);
                            tma_load_3d_pair(dst, &map_p0, 0, -1, m0, full(s));
                            tma_load_3d_pair(dst + kATile, &map_p0, p.Cin, -1, m0, full(s));
                        }
                    }
                    __syncwarp();
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
            } else
            for (long long tile = first_tile; tile < n_tiles && ok; tile += tile_stride) {
                const int nt = (int)(tile % p.tiles_n);
                const int m0 = (int)((tile / p.tiles_n) * per_pair + rank * (split ? p.mt + 1 : p.mt));
                const int m_mid = (int)((tile / p.tiles_n) * per_pair + p.mt);          // split: the cutout shared with the peer
                const int mid_row = rank == 0 ? 0 : p.split_h0;                         // ... and my first row of it
                const CUtensorMap* map_mid = rank == 0 ? &map_p0 : &map_p1;
                // the rows this CTA loads next (first column tile only: the others find them in L2 anyway)
                const long long tile_nx = tile + tile_stride;
                const bool warm = tile_nx < n_tiles && tile_nx % p.tiles_n == 0;
                const int m0_nx = (int)((tile_nx / p.tiles_n) * per_pair + rank * (split ? p.mt + 1 : p.mt));
                const int n0 = nt * BN + (int)rank * (BN / CG);          // my share of the weight tile's rows (CG = 1: all of them)
                int tap = 0, c0 = 0;
                for (int kb = 0; kb < n_kb; ++kb) {
                    if (!(ok = mbar_wait_warp(empty(s), ph ^ 1u, p.status, 1))) break;
                    const unsigned dst = ring + (unsigned)s * stage_bytes;
                    if (elect_one()) {
#ifndef POF_CONV_NO_PREFETCH
                        if (warm && tap == p.pad) {                      // the unshifted tap covers every row of the tile
                            tma_prefetch_3d(&map_a, c0, 0, m0_nx);
                            tma_prefetch_3d(&map_a, p.Cin + c0, 0, m0_nx);
                        }
#else
                        (void)warm; (void)m0_nx;
#endif
                        if (CG == 1) {
                            mbar_expect_tx(full(s), tx);
                            tma_load_3d(dst, &map_a, c0, tap - p.pad, m0, full(s));
                            tma_load_3d(dst + kATile, &map_a, p.Cin + c0, tap - p.pad, m0, full(s));
                            if (!p.w_resident) {
                                tma_load_2d(dst + 2 * kATile, &map_w, c0, (tap * 2 + 0) * p.Cout + n0, full(s));
                                tma_load_2d(dst + 2 * kATile + C::kBTile, &map_w, c0, (tap * 2 + 1) * p.Cout + n0, full(s));
                            }
                        } else {
                            if (rank == 0) mbar_expect_tx(full(s), tx);
                            tma_load_3d_pair(dst, &map_a, c0, tap - p.pad, m0, full(s));
                            tma_load_3d_pair(dst + kATile, &map_a, p.Cin + c0, tap - p.pad, m0, full(s));
                            if (split) {
                                const unsigned mid = dst + (unsigned)rows_full * kRowBytes;
                                tma_load_3d_pair(mid, map_mid, c0, mid_row + tap - p.pad, m_mid, full(s));
                                tma_load_3d_pair(mid + kATile, map_mid, p.Cin + c0, mid_row + tap - p.pad, m_mid, full(s));
                            }
                            if (!p.w_resident) {
                                tma_load_2d_pair(dst + 2 * kATile, &map_w, c0, (tap * 2 + 0) * p.Cout + n0, full(s));
                                tma_load_2d_pair(dst + 2 * kATile + C::kBTile, &map_w, c0, (tap * 2 + 1) * p.Cout + n0, full(s));
                            }
                        }
                    }
                    __syncwarp();
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                    c0 += kKBlock;
                    if (c0 == p.Cin) { c0 = 0; ++tap; }
                }
            }
        } else if (warp == 1 && rank == 0) {
            // ------------------------------------------------------------------ MMA issuer (the pair's leader only; whole warp, one elected lane issues)
            const unsigned idesc = umma_idesc(BN, kTileM * CG, F16);
            constexpr int kSteps = kRowBytes / 32;                          // one MMA consumes 32 B of the row: K = 8 (tf32) or 16 (f16)
            const unsigned long long desc_hi = umma_desc(0);                 // everything but the 14-bit start address
            unsigned ic = 0;                                               // chains issued so far
            int s = 0, in_chain = 0;
            unsigned ph = 0;
            bool ok = true;
            if (p.w_resident && first_tile < n_tiles) ok = mbar_wait_warp(wfull, 0u, p.status, 5);
            for (long long tile = first_tile; tile < n_tiles && ok; tile += tile_stride) {
                for (int kb = 0; kb < n_kb; ++kb) {
                    const unsigned buf = ic & 1u;
                    const bool first = in_chain == 0, last = in_chain + 1 == p.chain || kb + 1 == n_kb;
                    if (first && !(ok = mbar_wait_warp(tempty(buf), ((ic >> 1) & 1u) ^ 1u, p.status, 2))) break;   // chain ic-2 promoted
                    if ((!halo || kb == 0) && !(ok = mbar_wait_warp(full(s), ph, p.status, 3))) break;                // operands landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const bool stage_done = !halo || kb + 1 == n_kb;          // halo: the three taps (= k-blocks) read one stage
                    const unsigned st = ring + (unsigned)s * stage_bytes + (halo ? (unsigned)kb * kRowBytes : 0u);   // halo: the view of tap kb
                    const unsigned wst = p.w_resident ? base + (unsigned)kb * 2u * C::kBTile : st + 2 * kATile;
                    if (elect_one()) {
                        const unsigned long long a_hi = desc_hi | ((st >> 4) & 0x3fffu), a_lo = desc_hi | (((st + kATile) >> 4) & 0x3fffu);
                        const unsigned long long b_hi = desc_hi | ((wst >> 4) & 0x3fffu), b_lo = desc_hi | (((wst + C::kBTile) >> 4) & 0x3fffu);
                        const unsigned d = tmem_base + buf * BN;
                        auto mma = [&](unsigned long long a, unsigned long long b, unsigned acc) {
                            if (CG == 1) { if (F16) umma_f16(d, a, b, idesc, acc); else umma_tf32(d, a, b, idesc, acc); }
                            else { if (F16) umma_f16_pair(d, a, b, idesc, acc); else umma_tf32_pair(d, a, b, idesc, acc); }
                        };
#pragma unroll
                        for (int k = 0; k < kSteps; ++k) mma(a_lo + 2 * k, b_hi + 2 * k, k > 0 || !first);   // corrections first
#pragma unroll
                        for (int k = 0; k < kSteps; ++k) mma(a_hi + 2 * k, b_lo + 2 * k, 1);
#pragma unroll
                        for (int k = 0; k < kSteps; ++k) mma(a_hi + 2 * k, b_hi + 2 * k, 1);                 // main product last
                        if (CG == 1) {
                            if (stage_done) umma_commit(empty(s));
                            if (last) umma_commit(tfull(buf));
                        } else {
                            if (stage_done) umma_commit_pair(empty(s));      // both CTAs' producers may refill their slot
                            if (last) umma_commit_pair(tfull(buf));          // both CTAs' epilogues may read their rows
                        }
                    }
                    __syncwarp();
                    if (last) { ++ic; in_chain = 0; } else ++in_chain;
                    if (stage_done && ++s == n_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------------- promotion + epilogue
        if constexpr (C::kEpiWarps == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");   // 16 warps: compiled for <= 96 already
        const int e = warp - 4, q = e & 3, h = e >> 2;           // TMEM lane quarter (= warp % 4), column half
        const int row = q * 32 + lane;
        const unsigned lane_addr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(h * C::kAcc);
        long long it = 0;
        bool ok = true;
        for (long long tile = first_tile; tile < n_tiles && ok; tile += tile_stride) {
            float acc[C::kAcc];
#pragma unroll
            for (int j = 0; j < C::kAcc; ++j) acc[j] = 0.f;
            const int n_chains = (n_kb + p.chain - 1) / p.chain;
            for (int ch = 0; ch < n_chains; ++ch, ++it) {
                const int buf = (int)(it & 1);
                const unsigned bph = (unsigned)((it >> 1) & 1);
                if (!(ok = mbar_wait(tfull(buf), bph, p.status, 4))) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if constexpr (C::kAcc >= 64) {
#pragma unroll
                    for (int j = 0; j < C::kAcc / 64; ++j) {
                        unsigned v[64];
                        tmem_ld64(lane_addr + (unsigned)(buf * BN + j * 64), v);
#pragma unroll
                        for (int i = 0; i < 64; ++i) acc[j * 64 + i] = __fadd_rn(acc[j * 64 + i], __uint_as_float(v[i]));
                    }
                } else {
                    unsigned v[32];
                    tmem_ld32(lane_addr + (unsigned)(buf * BN), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = __fadd_rn(acc[i], __uint_as_float(v[i]));
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (CG == 1) mbar_arrive(tempty(buf));
                    else mbar_arrive_leader(tempty(buf));              // the leader's MMA warp waits for both CTAs' epilogues
                }
            }
            if (!ok) break;
            // epilogue for this tile: registers only, the tensor memory is already back with the MMA warp.
            // A lane holds one output ROW; written as is, every store instruction would touch 32 rows x 16 bytes.
            // Each warp therefore transposes 32 x 32 blocks through shared memory and stores 4 rows x 128 bytes
            // per instruction.
            const int nt = (int)(tile % p.tiles_n);
            // output row of lane 0, before pooling; lanes from `seg` on (the peer's part of a split cutout) are `seg_delta` rows away
            long long r0 = ((tile / p.tiles_n) * CG + rank) * rows_full + q * 32;
            int seg = 32, seg_delta = 0;
            if (split) {
                r0 = (tile / p.tiles_n) * per_pair * p.Lout + q * 32 + (rank == 0 ? 0 : kTileM + p.Lout - p.split_h0);
                if (rank != 0) {
                    seg = min(max(rows_full - q * 32, 0), 32);
                    seg_delta = -(rows_full + p.Lout - p.split_h0);
                }
            }
            bool in_tile = row < rows_tile;
            if (halo) {                                       // lane -> (cutout c, row l): at most one cutout boundary inside a warp (Lout + 2 >= 32)
                const int c_first = (q * 32) / halo_rows, c = row / halo_rows, l = row - c * halo_rows;
                r0 -= 2 * c_first;
                seg = min((c_first + 1) * halo_rows - q * 32, 32);
                seg_delta = -2;
                in_tile = c < p.mt && l < p.Lout;
            }
            const bool valid = in_tile && r0 + lane + (lane >= seg ? seg_delta : 0) < p.Mcut * p.Lout;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const int cbase = nt * BN + h * C::kAcc;
            float* stg = staging + e * (32 * kStagePitch);
            const int nrows = p.pool == 2 ? 16 : 32;                  // output rows this warp owns
            const long long orow0 = p.pool == 2 ? (r0 >> 1) : r0;
            const int oseg = p.pool == 2 ? seg >> 1 : seg, oseg_delta = p.pool == 2 ? seg_delta / 2 : seg_delta;
            // one transposition pass: lane l's 16 words (16 floats or 32 halves of ITS row) -> rows of 64 bytes at
            // dst + j * ld (bytes): four lanes store one output row, eight rows per instruction
            auto flush16 = [&](const unsigned* w, char* dst, long long ld) {
                unsigned* stw = reinterpret_cast<unsigned*>(stg);
#pragma unroll
                for (int k = 0; k < 16; k += 4)
                    *reinterpret_cast<uint4*>(stw + lane * kStagePitch + k) = make_uint4(w[k], w[k + 1], w[k + 2], w[k + 3]);
                __syncwarp();
                for (int j = lane >> 2; j < nrows; j += 8) {
                    const int src = p.pool == 2 ? 2 * j : j;          // the lane that holds output row j
                    if ((vmask >> src) & 1u)
                        st_stream_f4(reinterpret_cast<float4*>(dst + (j + (j >= oseg ? oseg_delta : 0)) * ld + (lane & 3) * 16),
                                     *reinterpret_cast<const float4*>(stw + src * kStagePitch + (lane & 3) * 4));
                }
                __syncwarp();
            };
            auto flush = [&](const float (&o)[32], float* dst, long long ld) {       // 32 fp32 columns
                unsigned w[16];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) w[k] = __float_as_uint(o[half * 16 + k]);
                    flush16(w, reinterpret_cast<char*>(dst + half * 16), ld * 4);
                }
            };
            auto flush_h = [&](const unsigned (&w)[16], __half* dst, long long ld) {  // 32 binary16 columns
                flush16(w, reinterpret_cast<char*>(dst), ld * 2);
            };
            float amax = 0.f;                                 // largest activation magnitude of this lane's row
#pragma unroll
            for (int g = 0; g < C::kAcc / 32; ++g) {
                float o[32];
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + cbase + g * 32 + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float v = acc[g * 32 + k + i];
                        if (p.pool == 2) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
                        // out_scale > 0 (the inverse of the weights' power of two, times the truncation compensation):
                        // the fused form rounds once, like a separate bias add, and it commutes with the max above
                        v = fmaf(v, p.out_scale, bb[i]);
                        o[k + i] = fmaxf(v, v * p.slope);               // LeakyReLU for 0 <= slope <= 1
                    }
                }
                if (p.out_plain) flush(o, p.out_plain + orow0 * p.Cout + cbase + g * 32, p.Cout);
                if (p.out_split) {
                    if (F16) {
                        __half* os = reinterpret_cast<__half*>(p.out_split) + orow0 * 2 * p.Cout + cbase + g * 32;
                        unsigned hi[16], lo[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const __half2 h2 = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
                            const float2 hf = __half22float2(h2);
                            const __half2 l2 = __floats2half2_rn(o[2 * k] - hf.x, o[2 * k + 1] - hf.y);
                            amax = fmaxf(amax, fmaxf(fabsf(o[2 * k]), fabsf(o[2 * k + 1])));
                            hi[k] = *reinterpret_cast<const unsigned*>(&h2);
                            lo[k] = *reinterpret_cast<const unsigned*>(&l2);
                        }
                        flush_h(hi, os, 2 * p.Cout);
                        flush_h(lo, os + p.Cout, 2 * p.Cout);
                    } else {
                        float* os = reinterpret_cast<float*>(p.out_split);
                        float hi[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) hi[k] = tf32_rn(o[k]);
                        flush(hi, os + orow0 * 2 * p.Cout + cbase + g * 32, 2 * p.Cout);
#pragma unroll
                        for (int k = 0; k < 32; ++k) o[k] = tf32_rn(o[k] - hi[k]);
                        flush(o, os + orow0 * 2 * p.Cout + p.Cout + cbase + g * 32, 2 * p.Cout);
                    }
                }
            }
            if (F16 && __any_sync(0xffffffffu, !(amax <= 65504.f) && valid) && lane == 0) atomicCAS(p.status, 0, 16);   // beyond binary16
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG > 1) cluster_sync();          // nobody leaves while the peer may still arrive on its barriers or read its operands
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

template <int BN, int CG, bool F16>
int launch(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mp0, const CUtensorMap& mp1, const Params& p, cudaStream_t stream) {
    using C = Cfg<BN, CG>;
    static bool attr_set[64] = {false};
    int dev = 0;
    POF_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        POF_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, CG, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
        attr_set[dev] = true;
    }
    const long long tiles = ((p.tiles_m + CG - 1) / CG) * p.tiles_n;
    const long long groups = sm_count() / CG;
    const int grid = CG * (int)(tiles < groups ? tiles : groups);
    conv_tc_kernel<BN, CG, F16><<<grid, C::kThreads, C::kSmem, stream>>>(ma, mw, mp0, mp1, p);
    POF_CUDA(cudaGetLastError());
    return POF_OK;
}

template <bool F16>
int launch_any(int bn, int cg, const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mp0, const CUtensorMap& mp1, const Params& p,
               cudaStream_t stream) {
    if (cg == 2) {
        if (bn == 256) return launch<256, 2, F16>(ma, mw, mp0, mp1, p, stream);
        if (bn == 128) return launch<128, 2, F16>(ma, mw, mp0, mp1, p, stream);
        return launch<64, 2, F16>(ma, mw, mp0, mp1, p, stream);
    }
    if (bn == 256) return launch<256, 1, F16>(ma, mw, mp0, mp1, p, stream);
    if (bn == 128) return launch<128, 1, F16>(ma, mw, mp0, mp1, p, stream);
    return launch<64, 1, F16>(ma, mw, mp0, mp1, p, stream);
}

// f16 = false: fp32 containers holding TF32 hi / lo parts; f16 = true: binary16 hi / lo parts.
int conv_tc_any(bool f16, const void* a_split, const void* w_split, const float* bias, long long Mcut, int LA, int Lout, int Cin,
                int Cout, int taps, int pad, int pool, float slope, float out_scale, float* out_plain, void* out_split, int* status,
                int chain_channels, cudaStream_t stream) {
    if (Mcut == 0) return POF_OK;
    const int kb = f16 ? KBlock<true>::value : KBlock<false>::value;
    const int esize = f16 ? 2 : 4;
    POF_REQUIRE(a_split && w_split && status && (out_plain || out_split), POF_ERR_NULL_POINTER, "pof_conv_tc_fwd: null pointer");
    POF_REQUIRE(Mcut > 0 && Mcut < (1ll << 31) && LA >= 1 && Lout >= 1 && Lout <= 128 && LA <= 256, POF_ERR_BAD_SHAPE,
                "pof_conv_tc_fwd: bad shape Mcut=%lld LA=%d Lout=%d", Mcut, LA, Lout);
    POF_REQUIRE(Cin >= kb && Cin % kb == 0, POF_ERR_BAD_SHAPE, "pof_conv_tc_fwd: Cin must be a multiple of %d (got %d)", kb, Cin);
    const int chain_flags = chain_channels;
    const int cg = (chain_channels & POF_CONV_TC_SINGLE_CTA) ? 1 : 2;      // high flag bits: tuning / tests only
    chain_channels &= ~(POF_CONV_TC_SINGLE_CTA | POF_CONV_TC_STREAM_W | POF_CONV_TC_NO_DEBIAS | POF_CONV_TC_NO_SPLIT_TILE | POF_CONV_TC_HALO);
    // binary16 chains have half as many accumulation steps per channel: 128 channels cost what 64 TF32 channels do
    // (6-7e-7 of the fp64 result per layer; cuDNN's fp32 kernels: 1-2e-6)
    if (chain_channels == 0) chain_channels = f16 ? 128 : 64;
    POF_REQUIRE(chain_channels > 0 && chain_channels % kb == 0, POF_ERR_BAD_PARAM,
                "pof_conv_tc_fwd: chain_channels must be a multiple of %d (got %d)", kb, chain_channels);
    POF_REQUIRE(Cout == 64 || Cout == 128 || (Cout >= 256 && Cout % 256 == 0), POF_ERR_BAD_SHAPE,
                "pof_conv_tc_fwd: Cout must be 64, 128 or a multiple of 256 (got %d)", Cout);
    POF_REQUIRE(taps >= 1 && pad >= 0 && pad < taps, POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: bad taps/pad %d/%d", taps, pad);
    POF_REQUIRE(pool == 1 || (pool == 2 && Lout % 2 == 0), POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: pool must be 1, or 2 with even Lout");
    POF_REQUIRE(out_scale > 0.f, POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: out_scale must be positive");
    POF_REQUIRE(slope >= 0.f && slope <= 1.f, POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: slope must lie in [0, 1] (got %g)", (double)slope);
    const uintptr_t al = reinterpret_cast<uintptr_t>(a_split) | reinterpret_cast<uintptr_t>(w_split) |
                         reinterpret_cast<uintptr_t>(out_plain) | reinterpret_cast<uintptr_t>(out_split) |
                         reinterpret_cast<uintptr_t>(bias);
    POF_REQUIRE((al & 15) == 0, POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: tensors must be 16-byte aligned");
    EncodeTiledFn enc = encode_tiled();
    POF_REQUIRE(enc != nullptr, POF_ERR_UNSUPPORTED, "pof_conv_tc_fwd: the driver does not export cuTensorMapEncodeTiled");

    const int bn = Cout >= 256 ? 256 : Cout;
    Params p;
    p.Mcut = Mcut;
    p.Lout = Lout;
    p.mt = kTileM / Lout;
    p.tiles_m = (Mcut + p.mt - 1) / p.mt;
    p.split_h0 = 0;
    if (cg == 2 && LA == Lout && !(chain_flags & POF_CONV_TC_NO_SPLIT_TILE) && (2 * p.mt + 1) * Lout <= 2 * kTileM && (kTileM - p.mt * Lout) % 2 == 0) {
        // one more cutout fits the pair's 256 rows than two separate tiles hold (Lout = 28: 9 instead of 8, 252 rows busy instead of 224)
        p.split_h0 = kTileM - p.mt * Lout;
        const long long per_pair = 2 * p.mt + 1;
        p.tiles_m = 2 * ((Mcut + per_pair - 1) / per_pair);          // the kernel counts work items as ceil(tiles_m / 2)
    }
    p.tiles_n = Cout / bn;
    p.Cin = Cin; p.Cout = Cout; p.taps = taps; p.pad = pad; p.pool = pool; p.slope = slope;
    p.chain = chain_channels / kb;
    p.out_scale = out_scale;
    if (!(chain_flags & POF_CONV_TC_NO_DEBIAS)) {
        // The tensor core adds each MMA's 16-deep sum to the running fp32 accumulator with TRUNCATION, so a chain's
        // partial sum comes out short: measured on every layer shape of the network (tools/conv_bias_probe.py,
        // profiles/r2_conv_truncation_bias.txt) the mean signed relative error of the outputs is -0.9e-7 / -2.8e-7 /
        // -6.8e-7 for chains of 4 / 8 / 16 main-product MMAs, i.e. -4.9e-8 per MMA beyond the first two, with a
        // spread (0.75e-7 / 1.6e-7 / 3.1e-7) smaller than the mean.  A one-sided error adds up coherently through the
        // eleven layers; the expected shortfall is therefore folded into the scale the epilogue applies anyway
        // (fma(sum, out_scale, bias): still one rounding).  What remains is zero-mean.
        // (TF32 parts: K = 8 per MMA, so a 64-channel chain has the 8 main MMAs of a 128-channel binary16 chain.)
        const float n_main = (float)chain_channels / (f16 ? 16.0f : 8.0f);
        p.out_scale = out_scale * (1.0f + 4.9e-8f * (n_main - 2.2f));
    }
    p.bias = bias; p.out_plain = out_plain; p.out_split = out_split; p.status = status;
    {   // narrow layers: this CTA's share of ALL weight k-blocks fits next to a deep ring of activation stages, so the
        // weights are loaded once per launch instead of once per tile (they are 35-55 % of those layers' L2->SM traffic)
        const int n_kb = taps * (Cin / kb);
        const int b_tile = (bn / cg) * kRowBytes;                       // one W part of one k-block
        const long long w_bytes = (long long)n_kb * 2 * b_tile;
        const int stage_a = 2 * kTileM * kRowBytes, stage_full = stage_a + 2 * b_tile;
        const bool forbid = (chain_flags & POF_CONV_TC_STREAM_W) != 0;
        const long long ring = (bn == 128 && POF_CONV_EPI16) ? kRingBytes - 8 * kStagingWarp : kRingBytes;      // Cfg<BN, CG>::kRing
        p.w_resident = !forbid && p.tiles_n == 1 && w_bytes <= ring - 4 * stage_a;
        const long long st = p.w_resident ? (ring - w_bytes) / stage_a : ring / stage_full;
        p.stages = (int)(st < kMaxStages ? st : kMaxStages);
    }
    // One load per tile instead of one per tap (see the kernel): a single k-block of channels, so the order of the MMAs and the
    // chain boundaries are those of the tap-by-tap form.  OPT-IN (POF_CONV_TC_HALO): it cuts the L2 -> shared-memory traffic of the
    // two 56-row layers by 2.9x and changes their time by -1.3 % / +1.5 % (64->128 0.844 -> 0.833 ms, 64->64 0.542 -> 0.550 ms at the
    // bench's launch size) - those layers are not waiting for the TMA (ncu: the stall samples sit in the epilogue warps) - and the wider layers would pay for the halo
    // rows with accumulator rows (126 -> 112 / 98 useful rows of 128 at 14 / 7 rows per cutout).
    p.halo = p.w_resident && !p.split_h0 && (chain_flags & POF_CONV_TC_HALO) && Cin == kb && taps == 3 && pad == 1 && LA == Lout &&
             Lout + 2 >= 32 && p.mt * (Lout + 2) <= kTileM;
    const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;

    alignas(64) CUtensorMap ma, mw, mp0, mp1;
    {   // A: [Mcut][LA][2 Cin], box (one 64-byte row of channels, Lout rows, mt cutouts); out-of-range rows read as zero.
        // Split tiles: two more boxes over the same tensor, one cutout high - the leader's and the peer's part of the shared cutout.
        const cuuint64_t dims[3] = {(cuuint64_t)(2 * Cin), (cuuint64_t)LA, (cuuint64_t)Mcut};
        const cuuint64_t strides[2] = {(cuuint64_t)(2 * Cin) * esize, (cuuint64_t)LA * (2 * Cin) * esize};
        const cuuint32_t es[3] = {1, 1, 1};
        CUtensorMap* maps[3] = {&ma, &mp0, &mp1};
        // Halo tiles: the second map's box is two rows higher (the padding rows of a cutout, zero filled).
        const cuuint32_t heights[3] = {(cuuint32_t)Lout, (cuuint32_t)(p.halo ? Lout + 2 : p.split_h0 ? p.split_h0 : Lout),
                                       (cuuint32_t)(p.split_h0 ? Lout - p.split_h0 : Lout)};
        for (int i = 0; i < 3; ++i) {
            const cuuint32_t box[3] = {(cuuint32_t)kb, heights[i], i == 0 || !p.split_h0 ? (cuuint32_t)p.mt : 1u};
            const CUresult r = enc(maps[i], dt, 3, const_cast<void*>(a_split), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, kRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, POF_CONV_L2_PROMO,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            POF_REQUIRE(r == CUDA_SUCCESS, POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
        }
    }
    {   // W: [taps][2][Cout][Cin] seen as a [taps * 2 * Cout, Cin] matrix, box (one 64-byte row of channels, bn rows)
        const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)taps * 2 * Cout};
        const cuuint64_t strides[1] = {(cuuint64_t)Cin * esize};
        const cuuint32_t box[2] = {(cuuint32_t)kb, (cuuint32_t)(bn / cg)};            // each CTA of a pair loads its half
        const cuuint32_t es[2] = {1, 1};
        const CUresult r = enc(&mw, dt, 2, const_cast<void*>(w_split), dims, strides, box, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, kRowBytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, POF_CONV_L2_PROMO,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        POF_REQUIRE(r == CUDA_SUCCESS, POF_ERR_BAD_PARAM, "pof_conv_tc_fwd: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    return f16 ? launch_any<true>(bn, cg, ma, mw, mp0, mp1, p, stream) : launch_any<false>(bn, cg, ma, mw, mp0, mp1, p, stream);
}

}  // namespace
}  // namespace pof

extern "C" {

int pof_conv_tc_fwd(const float* a_split, const float* w_split, const float* bias, long long Mcut, int LA, int Lout, int Cin,
                    int Cout, int taps, int pad, int pool, float slope, float* out_plain, float* out_split, int* status,
                    int chain_channels, void* stream_) {
    return pof::conv_tc_any(false, a_split, w_split, bias, Mcut, LA, Lout, Cin, Cout, taps, pad, pool, slope, 1.0f, out_plain,
                            out_split, status, chain_channels, (cudaStream_t)stream_);
}

int pof_conv_tc_f16_fwd(const void* a_split, const void* w_split, const float* bias, long long Mcut, int LA, int Lout, int Cin,
                        int Cout, int taps, int pad, int pool, float slope, float out_scale, float* out_plain, void* out_split,
                        int* status, int chain_channels, void* stream_) {
    return pof::conv_tc_any(true, a_split, w_split, bias, Mcut, LA, Lout, Cin, Cout, taps, pad, pool, slope, out_scale, out_plain,
                            out_split, status, chain_channels, (cudaStream_t)stream_);
}

}  // extern "C"
