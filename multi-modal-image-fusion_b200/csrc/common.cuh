// Common device/host helpers for libmmif_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mmif_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmmif_b200 is written for sm_100a (B200) only"
#endif

namespace mmif {

// ----------------------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define MMIF_CUDA(call)                                        \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return mmif::cuda_fail(e__, #call); \
    } while (0)

// Encode a 3-D (W,H,N) float32 tensor map with a (box_w, box_h, 1) box; false if TMA cannot be
// used for this tensor (row pitch or base not 16-byte aligned, driver entry point missing).
// Encodings are memoised per thread on (base, N, H, W, box): a training loop re-encodes nothing.
bool make_tensor_map(CUtensorMap* map, const float* base, int N, int H, int W, int box_w, int box_h);

// Launch bookkeeping behind mmif_launch_counts (include/mmif_b200.h MMIF_CNT_*): one relaxed atomic add per launch.
void count_launch(int which, unsigned long long n = 1ull);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute applies to the CURRENT device: one bit per device ordinal, so a process that drives several
// GPUs raises the dynamic shared-memory limit on each of them (a repeated set after a benign race is harmless).
static inline bool first_use_on_device(unsigned long long* mask) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    const unsigned long long bit = 1ull << (d & 63);
    // atomic test-and-set: two host threads racing on the first launch both see a consistent mask (a duplicate
    // cudaFuncSetAttribute would be harmless, a torn read-modify-write of the mask would not be)
    const unsigned long long prev = __atomic_fetch_or(mask, bit, __ATOMIC_ACQ_REL);
    return (prev & bit) == 0ull;
}

constexpr int kMaxWin = 17;
struct Taps {          // Gaussian taps of one window, passed by value in kernel parameter space
    float w[kMaxWin];  // float32 taps exactly as the reference builds them (loss.py:24-30)
    float wsum;        // sum of the reference's float32 2-D outer-product window (double -> float)
    float weps;        // wsum - 1 (computed in double)
    float wrho;        // wsum / (sum w)^2 - 1: the separable taps w_i*w_j (exact products) do not sum to
                       // what the reference's rounded products fl(w_i*w_j) sum to; ~1e-8, matters only where
                       // the variance of a flat region is compared with C2 ~ 1e-3
};
void make_taps(Taps* t, int win, double sigma);

#ifdef __CUDACC__
// ----------------------------------------------------------------------------- packed fp32x2 math
// FFMA2 / FMUL2 / FADD2 (sm_100): one issue slot for two fp32 lanes.  Measured on B200
// (tools/microbench/pipes.cu): FFMA 123 op/clk/SM, FFMA2 58 instr/clk/SM (= 117 fma/clk/SM), so
// packing does not raise the FMA roof but halves the issue slots the blur needs, leaving room for
// the LDS/ALU/MUFU work of the same warp.
#ifndef MMIF_SCALAR_FMA
#define MMIF_SCALAR_FMA 0
#endif
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
#if MMIF_SCALAR_FMA      // measurement aid: two scalar FFMA instead of one packed FFMA2 (tools/microbench/dispatch.cu: 2 x 1.03 vs 2.22 port cycles)
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bcast(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fmas(float w, float2 v, float2 acc) { return fma2(bcast(w), v, acc); }
__device__ __forceinline__ float2 muls(float w, float2 v) { return mul2(bcast(w), v); }
__device__ __forceinline__ float2 max2(float2 a, float b) { return make_float2(fmaxf(a.x, b), fmaxf(a.y, b)); }

// a/b with MUFU.RCP + one Newton step (<= 1 ulp for normal b); the SSIM quotients tolerate this
// (parity gate 1e-5 relative) and it keeps IEEE-division's ~10 instructions off the FMA pipe.
__device__ __forceinline__ float fdiv_nr(float a, float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    float e = fmaf(-b, r, 1.0f);
    r = fmaf(r, e, r);
    return a * r;
}
__device__ __forceinline__ float2 fdiv_nr2(float2 a, float2 b) { return f2(fdiv_nr(a.x, b.x), fdiv_nr(a.y, b.y)); }

// 1/b by MUFU.RCP alone (max relative error 2^-23): for quantities whose consumers tolerate one more ulp
__device__ __forceinline__ float rcp_approx(float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return r;
}
__device__ __forceinline__ float2 rcp2(float2 b) { return f2(rcp_approx(b.x), rcp_approx(b.y)); }

__device__ __forceinline__ float finite_or_zero(float v) { return (fabsf(v) <= 3.0e38f) ? v : 0.0f; }

// ----------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum NV doubles per thread over a block of NTHREADS (multiple of 32); result valid in thread 0.
// `scratch` needs NV * (NTHREADS/32) doubles. Deterministic (fixed tree).
template <int NV, int NTHREADS>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = NTHREADS / 32;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = warp_sum(v[i]);
        if (lane == 0) scratch[i * NW + warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < NW; ++w) s += scratch[i * NW + w];
            v[i] = s;
        }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------- mbarrier + TMA
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load (UTMALDG): box lands densely at `dst`, completion bytes on `bar`.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif  // __CUDACC__

}  // namespace mmif
