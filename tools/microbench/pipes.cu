// Micro-benchmarks that back the design choices in DESIGN.md (sm_100a).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Measures issue throughput of the instructions the stencil kernels lean on:
// FFMA (3-reg / const operand), FFMA2 (fma.rn.f32x2), MUFU.RCP, full fp32 division,
// DFMA, shared-memory and L2 atomics at several bin spreads, __match_any_sync.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
__constant__ float cw[16];

__global__ void k_ffma(float* out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float x = a + threadIdx.x, y = b;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// accumulate form: acc += w*v with distinct v regs (as in the blur)
__global__ void k_ffma_acc(float* out, const float* in) {
    float acc[16], v[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x + i * 32];
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(cw[(i + k) & 15], v[k], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += 1.0f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void fma2(float2& d, const float2& a, const float2& b) {
    asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\t"
                 "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rd, {%0, %1};\n\t"
                 "fma.rn.f32x2 rd, ra, rb, rd;\n\tmov.b64 {%0, %1}, rd;\n\t}"
                 : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

__global__ void k_ffma2(float* out, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i);
    float2 x = make_float2(a + threadIdx.x, a), y = make_float2(b, b + 1);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            // acc = acc*x + y  (d = a*b + c form needs c=y: emulate as acc = fma(acc,x,y))
            asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\t"
                         "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rc, {%4, %5};\n\t"
                         "fma.rn.f32x2 ra, ra, rb, rc;\n\tmov.b64 {%0, %1}, ra;\n\t}"
                         : "+f"(acc[i].x), "+f"(acc[i].y) : "f"(x.x), "f"(x.y), "f"(y.x), "f"(y.y));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// blur-like: acc2[i] += w2[k]*v2[k]
__global__ void k_ffma2_acc(float* out, const float* in) {
    float2 acc[16], v[8], w[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = make_float2(in[threadIdx.x + i * 32], in[threadIdx.x + i * 32 + 1]); w[i] = make_float2(in[i], in[i]); }
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < 16; ++i) fma2(acc[i], w[(i + k) & 7], v[k]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i].x += 1.0f; v[i].y += 1.0f; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_rcp(float* out, float a) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __frcp_rn(acc[i]) + 1.5f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rcp_approx(float* out, float a) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r + 1.5f; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_div(float* out, float a) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = __fdiv_rn(a, acc[i]) + 1.5f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// LDS.64 + FFMA2 mix (1 LDS.64 per 4 FFMA2), like the vertical pass
__global__ void k_lds_ffma2(float* out, float a) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i * a, i);
    __syncthreads();
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
    float2 w = make_float2(a, a);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 v = sm[(idx + k * 256) & 2047];
#pragma unroll
            for (int i = 0; i < 4; ++i) fma2(acc[k * 4 + i], w, v);
        }
        idx += 7;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// shared atomics into `bins` bins (random), 256 threads
__global__ void k_atoms(unsigned* out, int bins, int spread_mode) {
    extern __shared__ unsigned h[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        uint32_t r = rng(s);
        int b = spread_mode == 0 ? (r % bins) : (spread_mode == 1 ? ((it >> 4) % bins) : ((threadIdx.x >> 3) + (r & 3)) % bins);
        atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    unsigned t = 0;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) t += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// same with match_any aggregation
__global__ void k_atoms_match(unsigned* out, int bins, int spread_mode) {
    extern __shared__ unsigned h[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < ITERS; ++it) {
        uint32_t r = rng(s);
        int b = spread_mode == 0 ? (r % bins) : (spread_mode == 1 ? ((it >> 4) % bins) : ((threadIdx.x >> 3) + (r & 3)) % bins);
        unsigned m = __match_any_sync(0xffffffffu, b);
        if (lane == (__ffs(m) - 1)) atomicAdd(&h[b], (unsigned)__popc(m));
    }
    __syncthreads();
    unsigned t = 0;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) t += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// global RED into 65536 bins
__global__ void k_redg(unsigned* hist, int bins, int spread_mode) {
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 977u;
    for (int it = 0; it < ITERS / 4; ++it) {
        uint32_t r = rng(s);
        int b = spread_mode == 0 ? (r % bins) : (spread_mode == 1 ? ((it >> 4) % bins) : (((r & 255) * 257 + ((r >> 8) & 3)) % bins));
        atomicAdd(&hist[b], 1u);
    }
}

template <typename F>
static float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms=%d maxclk=%d MHz\n", p.name, sms, clk_khz / 1000);
    float* out; CK(cudaMalloc(&out, 64 << 20)); CK(cudaMemset(out, 0, 64 << 20));
    float* in; CK(cudaMalloc(&in, 1 << 20)); CK(cudaMemset(in, 0, 1 << 20));
    float hw[16]; for (int i = 0; i < 16; ++i) hw[i] = 0.01f * i; CK(cudaMemcpyToSymbol(cw, hw, sizeof(hw)));
    const int blocks = sms * 8, threads = 256;
    auto report = [&](const char* name, float ms, double ops_per_thread) {
        double total = ops_per_thread * blocks * (double)threads;
        printf("%-28s %8.3f ms  %8.2f Gop/s  %7.2f op/clk/SM @maxclk\n", name, ms, total / ms * 1e-6, total / (ms * 1e-3) / sms / (clk_khz * 1e3));
    };
    report("FFMA chain(16) 3-reg", timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
    report("FFMA acc+=c[w]*v", timeit([&] { k_ffma_acc<<<blocks, threads>>>(out, in); }), 16.0 * ITERS);
    report("FFMA2 chain(16) (x2 flops)", timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16.0 * ITERS);
    report("FFMA2 acc+=w*v", timeit([&] { k_ffma2_acc<<<blocks, threads>>>(out, in); }), 16.0 * ITERS);
    report("LDS.64+4xFFMA2 (count FFMA2)", timeit([&] { k_lds_ffma2<<<blocks, threads>>>(out, 1.0f); }), 4.0 * ITERS);
    report("rcp_rn", timeit([&] { k_rcp<<<blocks, threads>>>(out, 1.5f); }), 8.0 * ITERS);
    report("rcp.approx", timeit([&] { k_rcp_approx<<<blocks, threads>>>(out, 1.5f); }), 8.0 * ITERS);
    report("div_rn", timeit([&] { k_div<<<blocks, threads>>>(out, 1.5f); }), 8.0 * ITERS);
    report("DFMA chain(8)", timeit([&] { k_dfma<<<blocks, threads>>>((double*)out, 1.0001, 0.5); }), 8.0 * ITERS);
    const char* modes[3] = {"random", "uniform-warp", "clustered"};
    for (int bins : {256, 4096, 16384}) for (int m = 0; m < 3; ++m) {
        char nm[64]; snprintf(nm, 64, "ATOMS %5d %s", bins, modes[m]);
        CK(cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_atoms_match, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        report(nm, timeit([&] { k_atoms<<<blocks, threads, bins * 4>>>((unsigned*)out, bins, m); }), (double)ITERS);
        snprintf(nm, 64, "ATOMS+match %5d %s", bins, modes[m]);
        report(nm, timeit([&] { k_atoms_match<<<blocks, threads, bins * 4>>>((unsigned*)out, bins, m); }), (double)ITERS);
    }
    unsigned* gh; CK(cudaMalloc(&gh, 65536 * 4)); CK(cudaMemset(gh, 0, 65536 * 4));
    for (int m = 0; m < 3; ++m) {
        char nm[64]; snprintf(nm, 64, "REDG 65536 %s", modes[m]);
        report(nm, timeit([&] { k_redg<<<blocks, threads>>>(gh, 65536, m); }), ITERS / 4.0);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
