// Can the idle FP64 pipe take blur work off the FP32 FMA pipe?  (sm_100a)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coissue coissue.cu
// Per loop iteration every thread issues NF independent FFMA2 (acc += w*v, the blur form) and ND
// independent DFMA; if the two pipes run concurrently the time of (NF, ND) equals
// max(time(NF, 0), time(0, ND)) as long as the issue slots suffice.  Also: F2F.F64.F32 / F2F.F32.F64
// conversion throughput (the price of feeding fp32 data to the FP64 pipe), and the SM clock / power
// under the mixed load (a clock drop would cancel the gain).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
constexpr int ITERS = 8192;

__device__ __forceinline__ void fma2(float2& d, const float2& a, const float2& b) {
    asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\t"
                 "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rd, {%0, %1};\n\t"
                 "fma.rn.f32x2 rd, ra, rb, rd;\n\tmov.b64 {%0, %1}, rd;\n\t}"
                 : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

template <int NF, int ND>
__global__ void __launch_bounds__(128) k_mix(float* out, const float* in) {
    float2 facc[NF > 0 ? NF : 1];
    double dacc[ND > 0 ? ND : 1];
    float2 w = make_float2(in[0], in[1]), v = make_float2(in[threadIdx.x], in[threadIdx.x + 1]);
    double dw = in[2], dv = in[threadIdx.x + 3];
#pragma unroll
    for (int i = 0; i < NF; ++i) facc[i] = make_float2(i, i);
#pragma unroll
    for (int i = 0; i < ND; ++i) dacc[i] = i;
    for (int it = 0; it < ITERS; ++it) {
        // interleave the two streams in program order
#pragma unroll
        for (int i = 0; i < (NF > ND ? NF : ND); ++i) {
            if (i < NF) fma2(facc[i], w, v);
            if (i < ND) dacc[i] = fma(dw, dv, dacc[i]);
        }
        v.x += 1.f; dv += 1.0;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += facc[i].x + facc[i].y;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += (float)dacc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// blur-like DFMA with fp32 inputs: 1 conversion feeds 8 DFMA (H-pass shape)
template <int CVT_PER_8>
__global__ void __launch_bounds__(128) k_cvt_dfma(float* out, const float* in) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = i;
    float v = in[threadIdx.x];
    const double w = in[1];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CVT_PER_8; ++c) {
            const double dv = (double)(v + (float)c);
#pragma unroll
            for (int i = c; i < 8; i += CVT_PER_8) acc[i] = fma(w, dv, acc[i]);
        }
        v += 1.f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (float)acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(128) k_cvt(float* out, const float* in) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x + i];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { double d = (double)v[i]; asm volatile("" : "+d"(d)); v[i] = (float)d; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    const int blocks = sms * 8, threads = 128;      // 8 warps / SM resident at a time per wave of 2 CTAs? (grid >> SMs)
    auto rep = [&](const char* name, float ms, double nf, double nd) {
        const double thr = (double)blocks * threads * ITERS;
        const double cyc = ms * 1e-3 * clk_khz * 1e3;       // cycles at max clock
        printf("%-34s %8.3f ms   FFMA2 %6.2f /clk/SM   DFMA %6.2f /clk/SM   (fp32-FMA-equivalent %6.1f /clk/SM)\n", name, ms,
               nf * thr / cyc / sms, nd * thr / cyc / sms, (2 * nf + nd) * thr / cyc / sms);
    };
#define RUN(NF, ND) rep("mix NF=" #NF " ND=" #ND, timeit([&] { k_mix<NF, ND><<<blocks, threads>>>(out, in); }), NF, ND)
    RUN(16, 0); RUN(0, 16); RUN(0, 8);
    RUN(16, 4); RUN(16, 8); RUN(16, 12); RUN(16, 16); RUN(12, 16); RUN(8, 16);
    RUN(24, 8); RUN(24, 12);
    {
        const double thr = (double)blocks * threads * ITERS;
        float ms = timeit([&] { k_cvt<<<blocks, threads>>>(out, in); });
        printf("F2F f32->f64->f32 pairs            %8.3f ms   %6.2f pairs/clk/SM\n", ms, 8 * thr / (ms * 1e-3 * clk_khz * 1e3) / sms);
        ms = timeit([&] { k_cvt_dfma<1><<<blocks, threads>>>(out, in); });
        printf("1 cvt + 8 DFMA                     %8.3f ms   DFMA %6.2f /clk/SM\n", ms, 8 * thr / (ms * 1e-3 * clk_khz * 1e3) / sms);
        ms = timeit([&] { k_cvt_dfma<2><<<blocks, threads>>>(out, in); });
        printf("2 cvt + 8 DFMA                     %8.3f ms   DFMA %6.2f /clk/SM\n", ms, 8 * thr / (ms * 1e-3 * clk_khz * 1e3) / sms);
    }
    // sustained mixed load for the clock / power sample taken by the caller (nvidia-smi in parallel)
    for (int r = 0; r < 40; ++r) k_mix<16, 8><<<blocks * 8, threads>>>(out, in);
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
