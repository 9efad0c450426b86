// How many scheduler (dispatch-port) cycles does a packed FFMA2 cost, and can another instruction use its second cycle?  (sm_100a)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dispatch dispatch.cu
// Loop body = NF independent FFMA2 (or 2*NF scalar FFMA when PACK == 0) interleaved with NA independent LOP3 (ALU pipe) at 1, 2
// and 4 warps per scheduler; reported: cycles per loop iteration per scheduler.  If the packed instruction held only the FMA
// pipe for its second cycle, NA <= NF LOP3 would be free; if it holds the dispatch port, every LOP3 adds a cycle.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void fma2(float2& d, const float2& a, const float2& b) {
    asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\t"
                 "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rd, {%0, %1};\n\t"
                 "fma.rn.f32x2 rd, ra, rb, rd;\n\tmov.b64 {%0, %1}, rd;\n\t}"
                 : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

template <int NF, int NA, int PACK>
__global__ void __launch_bounds__(128) k_mix(float* out, const float* in, int iters) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(i, i);
    float2 w = make_float2(in[0], in[1]), v = make_float2(in[threadIdx.x], in[threadIdx.x + 1]);
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i < NF) {
                    if (PACK) fma2(acc[i], w, v);
                    else { acc[i].x = fmaf(w.x, v.x, acc[i].x); acc[i].y = fmaf(w.y, v.y, acc[i].y); }
                }
                if (i < NA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i & 7]) : "r"(u), "r"(i));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (float)x[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NA, int PACK>
static int run(float* out, float* in, int sms, int clk_khz, int ctas_per_sm) {
    const int iters = 20000;
    const size_t smem = ctas_per_sm == 1 ? 120 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : 50 * 1024);   // pins the occupancy
    CK(cudaFuncSetAttribute(k_mix<NF, NA, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    const int blocks = sms * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<NF, NA, PACK><<<blocks, 128, smem>>>(out, in, iters); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_mix<NF, NA, PACK><<<blocks, 128, smem>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // one warp per scheduler per CTA: a scheduler runs ctas_per_sm warps; cycles per (4 x body) per scheduler
    const double cyc = best * 1e-3 * clk_khz * 1e3 / ((double)iters * 4 * ctas_per_sm);
    printf("%s x%-2d + LOP3 x%-2d  %d warp(s)/scheduler: %7.2f cycles per body\n", PACK ? "FFMA2" : "FFMA ", PACK ? NF : 2 * NF, NA, ctas_per_sm, cyc);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms=%d maxclk=%d MHz\n", p.name, sms, clk_khz / 1000);
    float* out; CK(cudaMalloc(&out, 64 << 20)); CK(cudaMemset(out, 0, 64 << 20));
    float* in; CK(cudaMalloc(&in, 1 << 20)); CK(cudaMemset(in, 0, 1 << 20));
    for (int c = 1; c <= 4; c *= 2) {
        run<16, 0, 1>(out, in, sms, clk_khz, c); run<16, 4, 1>(out, in, sms, clk_khz, c); run<16, 8, 1>(out, in, sms, clk_khz, c);
        run<16, 16, 1>(out, in, sms, clk_khz, c); run<8, 16, 1>(out, in, sms, clk_khz, c); run<0, 16, 1>(out, in, sms, clk_khz, c);
        run<16, 0, 0>(out, in, sms, clk_khz, c); run<16, 8, 0>(out, in, sms, clk_khz, c); run<16, 16, 0>(out, in, sms, clk_khz, c);
    }
    printf("done\n");
    return 0;
}
