// Small data-movement operators around the stencil kernels: the MS-SSIM level step of the loss
// (loss.py:147-153) and its adjoint, reflect padding (use_padding=True, loss.py:45-47) and its adjoint,
// and the TVLoss backward (loss.py:347-358).
#include "common.cuh"

namespace mmif {

__global__ void __launch_bounds__(256) halve1_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Ho, int Wo) {
    const int n = blockIdx.z;
    const int j = blockIdx.x * 64 + (threadIdx.x & 63), i = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= Ho || j >= Wo) return;
    const float* s = src + (size_t)n * H * W;
    const int r0 = 2 * i, r1 = (2 * i + 1 < H) ? 2 * i + 1 : H - 2;
    const int c0 = 2 * j, c1 = (2 * j + 1 < W) ? 2 * j + 1 : W - 2;
    const float v = ((__ldg(s + (size_t)r0 * W + c0) + __ldg(s + (size_t)r0 * W + c1)) + __ldg(s + (size_t)r1 * W + c0)) +
                    __ldg(s + (size_t)r1 * W + c1);
    dst[((size_t)n * Ho + i) * Wo + j] = v * 0.25f;
}

// g_src += adjoint(halve)(g_dst): every source pixel belongs to one 2x2 cell; with an odd size the
// reflected pad row/column H (W) is a second reader of row H-2 (column W-2).
__global__ void __launch_bounds__(256) halve1_bwd_kernel(const float* __restrict__ gd, float* __restrict__ gs, int H, int W, int Ho, int Wo) {
    const int n = blockIdx.z;
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= H || c >= W) return;
    const float* g = gd + (size_t)n * Ho * Wo;
    const bool er = (H & 1) && (r == H - 2), ec = (W & 1) && (c == W - 2);
    float v = __ldg(g + (size_t)(r >> 1) * Wo + (c >> 1));
    if (er) v += __ldg(g + (size_t)((H - 1) >> 1) * Wo + (c >> 1));
    if (ec) v += __ldg(g + (size_t)(r >> 1) * Wo + ((W - 1) >> 1));
    if (er && ec) v += __ldg(g + (size_t)((H - 1) >> 1) * Wo + ((W - 1) >> 1));
    gs[((size_t)n * H + r) * W + c] += 0.25f * v;
}

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(256) reflect_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int pad) {
    const int n = blockIdx.z, Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= Hp || c >= Wp) return;
    dst[((size_t)n * Hp + r) * Wp + c] = __ldg(src + ((size_t)n * H + reflect_idx(r - pad, H)) * W + reflect_idx(c - pad, W));
}

// g_src = adjoint(reflect_pad)(g_dst): gather over the (up to 3 x 3) padded positions that read pixel (r, c).
__global__ void __launch_bounds__(256) reflect_pad_bwd_kernel(const float* __restrict__ gd, float* __restrict__ gs, int H, int W, int pad) {
    const int n = blockIdx.z, Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= H || c >= W) return;
    const float* g = gd + (size_t)n * Hp * Wp;
    int rs[3] = {r, -r, 2 * H - 2 - r}, cs[3] = {c, -c, 2 * W - 2 - c};     // unpadded coordinates of the readers
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int rr = rs[a];
        if (rr < -pad || rr >= H + pad || (a > 0 && rr == r)) continue;       // r == 0 / H-1 reflect onto themselves: count once
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int cc = cs[b];
            if (cc < -pad || cc >= W + pad || (b > 0 && cc == c)) continue;
            v += __ldg(g + (size_t)(rr + pad) * Wp + (cc + pad));
        }
    }
    gs[((size_t)n * H + r) * W + c] = v;
}

__device__ __forceinline__ float nder(float d, int norm) { return norm == MMIF_NORM_L1 ? ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f)) : 2.f * d; }

__global__ void __launch_bounds__(256) tv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gout, float* __restrict__ gx,
                                                     int H, int W, int norm, float kv, float kh) {
    const int n = blockIdx.z;
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= H || c >= W) return;
    const float* p = x + (size_t)n * H * W;
    const float g = __ldg(gout), v = __ldg(p + (size_t)r * W + c);
    float acc = 0.f;
    if (r > 0) acc += kv * nder(v - __ldg(p + (size_t)(r - 1) * W + c), norm);
    if (r + 1 < H) acc -= kv * nder(__ldg(p + (size_t)(r + 1) * W + c) - v, norm);
    if (c > 0) acc += kh * nder(v - __ldg(p + (size_t)r * W + c - 1), norm);
    if (c + 1 < W) acc -= kh * nder(__ldg(p + (size_t)r * W + c + 1) - v, norm);
    gx[((size_t)n * H + r) * W + c] = g * acc;
}

// ---- uint8 ingest (eval.py:182-194 decodes 8-bit images and widens them on the host; widening on the device
// cuts the host->device traffic of an evaluation 4x).  16 pixels per thread: one 16-byte load, four 16-byte stores.
// UNIT: divide by 255 in IEEE float32 division — bit-identical to the `uint8 -> float32 / 255` of the reference's input
// pipeline (data/dataset.py through torchvision's to_tensor), so a training loop can ship 8-bit sources to the device.
template <bool UNIT>
__device__ __forceinline__ float widen1(unsigned v) { return UNIT ? __fdiv_rn((float)v, 255.0f) : (float)v; }

template <bool UNIT>
__global__ void __launch_bounds__(256) widen_u8_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, size_t n, int vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const size_t n16 = n >> 4;
        const uint4* s16 = reinterpret_cast<const uint4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (; i < n16; i += stride) {
            const uint4 v = __ldg(s16 + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                d4[4 * i + k] = make_float4(widen1<UNIT>(w[k] & 0xFF), widen1<UNIT>((w[k] >> 8) & 0xFF), widen1<UNIT>((w[k] >> 16) & 0xFF),
                                            widen1<UNIT>(w[k] >> 24));
        }
        for (size_t t = (n16 << 4) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) dst[t] = widen1<UNIT>(src[t]);
    } else {
        for (; i < n; i += stride) dst[i] = widen1<UNIT>(src[i]);
    }
}

// ---- NormLoss (loss.py:361-385): weight * mean(|x|) or weight * mean(x^2) of an arbitrary tensor, and its gradient.
// Deterministic: per-CTA partials in double, the last CTA to arrive adds them in index order.
constexpr int kNormBlocks = 592;
__global__ void __launch_bounds__(256) norm_loss_kernel(const float* __restrict__ x, size_t n, int norm, double scale, unsigned* counter,
                                                        double* partial, double* out) {
    __shared__ double red[8];
    __shared__ int flag;
    float acc = 0.f;
    double total = 0.0;
    int cnt = 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = __ldg(x + i);
        acc += (norm == MMIF_NORM_L1) ? fabsf(v) : v * v;
        if (++cnt == 64) { total += (double)acc; acc = 0.f; cnt = 0; }       // bound the float run length
    }
    double v1[1] = {total + (double)acc};
    block_sum<1, 256>(v1, red);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = v1[0];
        __threadfence();
        flag = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!flag) return;
    __threadfence();
    double t[1] = {0.0};
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) t[0] += __ldcg(partial + k);
    block_sum<1, 256>(t, red);
    if (threadIdx.x == 0) { out[0] = t[0] * scale; *counter = 0u; }
}
__global__ void __launch_bounds__(256) norm_loss_bwd_kernel(const float* __restrict__ x, size_t n, int norm, float k, const float* __restrict__ gout,
                                                            float* __restrict__ gx) {
    const float g = __ldg(gout) * k;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = __ldg(x + i);
        gx[i] = g * ((norm == MMIF_NORM_L1) ? ((v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f)) : 2.f * v);
    }
}

static int chk(const void* a, const void* b, int N, int H, int W) {
    if (!a || !b) { set_error("null pointer"); return MMIF_E_NULL; }
    if (N < 1 || H < 2 || W < 2) { set_error("bad shape (%d,%d,%d)", N, H, W); return MMIF_E_SHAPE; }
    if (((uintptr_t)a | (uintptr_t)b) & 3) { set_error("pointers must be 4-byte aligned"); return MMIF_E_ALIGN; }
    return MMIF_OK;
}

}  // namespace mmif

using namespace mmif;

extern "C" int mmif_halve(const float* src, int N, int H, int W, float* dst, void* stream) {
    int rc = chk(src, dst, N, H, W); if (rc) return rc;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    halve1_kernel<<<dim3(ceil_div(Wo, 64), ceil_div(Ho, 4), N), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, Ho, Wo);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
extern "C" int mmif_halve_bwd(const float* g_dst, int N, int H, int W, float* g_src_accum, void* stream) {
    int rc = chk(g_dst, g_src_accum, N, H, W); if (rc) return rc;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    halve1_bwd_kernel<<<dim3(ceil_div(W, 64), ceil_div(H, 4), N), 256, 0, (cudaStream_t)stream>>>(g_dst, g_src_accum, H, W, Ho, Wo);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
extern "C" int mmif_reflect_pad(const float* src, int N, int H, int W, int pad, float* dst, void* stream) {
    int rc = chk(src, dst, N, H, W); if (rc) return rc;
    if (pad < 0 || pad >= H || pad >= W) { set_error("reflect pad %d must be smaller than H, W", pad); return MMIF_E_SHAPE; }
    reflect_pad_kernel<<<dim3(ceil_div(W + 2 * pad, 64), ceil_div(H + 2 * pad, 4), N), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, pad);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
extern "C" int mmif_reflect_pad_bwd(const float* g_dst, int N, int H, int W, int pad, float* g_src, void* stream) {
    int rc = chk(g_dst, g_src, N, H, W); if (rc) return rc;
    if (pad < 0 || pad >= H || pad >= W) { set_error("reflect pad %d must be smaller than H, W", pad); return MMIF_E_SHAPE; }
    reflect_pad_bwd_kernel<<<dim3(ceil_div(W, 64), ceil_div(H, 4), N), 256, 0, (cudaStream_t)stream>>>(g_dst, g_src, H, W, pad);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
extern "C" int mmif_tv_loss_bwd(const float* x, int N, int H, int W, int norm, float weight, const float* gout1, float* gx,
                                void* stream) {
    int rc = chk(x, gx, N, H, W); if (rc) return rc;
    if (!gout1) { set_error("null gout1"); return MMIF_E_NULL; }
    if (norm != MMIF_NORM_L1 && norm != MMIF_NORM_L2) { set_error("unsupported norm"); return MMIF_E_MODE; }
    const float kv = weight / ((float)N * (float)(H - 1) * (float)W), kh = weight / ((float)N * (float)H * (float)(W - 1));
    tv_bwd_kernel<<<dim3(ceil_div(W, 64), ceil_div(H, 4), N), 256, 0, (cudaStream_t)stream>>>(x, gout1, gx, H, W, norm, kv, kh);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}

static int widen_impl(const unsigned char* src, size_t n, float* dst, void* stream, bool unit);
extern "C" int mmif_widen_u8(const unsigned char* src, size_t n, float* dst, void* stream) { return widen_impl(src, n, dst, stream, false); }
extern "C" int mmif_widen_u8_unit(const unsigned char* src, size_t n, float* dst, void* stream) { return widen_impl(src, n, dst, stream, true); }
static int widen_impl(const unsigned char* src, size_t n, float* dst, void* stream, bool unit) {
    if (!src || !dst) { set_error("null pointer"); return MMIF_E_NULL; }
    if (((uintptr_t)dst) & 3) { set_error("dst must be 4-byte aligned"); return MMIF_E_ALIGN; }
    if (n == 0) return MMIF_OK;
    const int vec = ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0;
    size_t blocks = (n / 16 + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks);
    if (unit) widen_u8_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n, vec);
    else widen_u8_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n, vec);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
extern "C" size_t mmif_norm_workspace_bytes(void) { return 256 + (size_t)kNormBlocks * sizeof(double); }
extern "C" int mmif_norm_loss(const float* x, size_t n, int norm, float weight, double* out, void* ws, size_t ws_bytes, void* stream) {
    if (!x || !out || !ws) { set_error("null pointer"); return MMIF_E_NULL; }
    if (n == 0) { set_error("empty tensor"); return MMIF_E_SHAPE; }
    if (norm != MMIF_NORM_L1 && norm != MMIF_NORM_L2) { set_error("unsupported norm"); return MMIF_E_MODE; }
    if (ws_bytes < mmif_norm_workspace_bytes() || (((uintptr_t)ws) & 7)) { set_error("norm workspace too small / unaligned"); return MMIF_E_WORKSPACE; }
    size_t blocks = (n + 2047) / 2048;
    blocks = blocks < 1 ? 1 : (blocks > (size_t)kNormBlocks ? (size_t)kNormBlocks : blocks);
    norm_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, norm, (double)weight / (double)n, (unsigned*)ws,
                                                                          (double*)((unsigned char*)ws + 256), out);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
extern "C" int mmif_norm_loss_bwd(const float* x, size_t n, int norm, float weight, const float* gout1, float* gx, void* stream) {
    if (!x || !gout1 || !gx) { set_error("null pointer"); return MMIF_E_NULL; }
    if (norm != MMIF_NORM_L1 && norm != MMIF_NORM_L2) { set_error("unsupported norm"); return MMIF_E_MODE; }
    if (n == 0) return MMIF_OK;
    size_t blocks = (n + 1023) / 1024;
    blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
    norm_loss_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, norm, weight / (float)n, gout1, gx);
    mmif::count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
