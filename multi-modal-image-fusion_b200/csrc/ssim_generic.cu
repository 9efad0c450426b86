// Generic SSIM of the loss module (reference core/loss.py:52-110) for ANY window size 1..17, odd or even, with every
// gradient the reference's autograd provides: per-sample means or per-position maps of ssim / cs / sigma, and the
// backward with respect to BOTH images for upstream gradients given per sample or per position (size_average=False).
// This is the slow-but-complete companion of the strip kernels (which are built for the windows 11/9/7/5/3 and for the
// gradients the training path needs): one thread per window position evaluates the k x k float32 window
// W[i][j] = fl(w_i * w_j) directly (exactly the torch.mm window of loss.py:36-37) with double accumulation, so its values
// sit between the reference's float32 and float64 evaluations; the backward is a coefficient pass + a k x k gather.
#include "metrics.cuh"

namespace mmif {

struct GenMoments { double mx, my, exx, eyy, exy; };

__device__ __forceinline__ GenMoments gen_moments(const float* __restrict__ x, const float* __restrict__ y, int W, int i0, int j0, int k,
                                                  const Taps& taps) {
    GenMoments m = {0, 0, 0, 0, 0};
    for (int u = 0; u < k; ++u)
        for (int v = 0; v < k; ++v) {
            const double w = (double)(float)(taps.w[u] * taps.w[v]);
            const size_t p = (size_t)(i0 + u) * W + (j0 + v);
            const double a = x[p], b = y[p];
            m.mx += w * a; m.my += w * b; m.exx += w * a * a; m.eyy += w * b * b; m.exy += w * a * b;
        }
    return m;
}

// forward: per-position ssim / cs / sigma (optionally stored as maps) and their per-sample sums
__global__ void __launch_bounds__(256)
ssim_generic_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, int H, int W, int k, const Taps taps, double C1, double C2,
                        float* map_ssim, float* map_cs, float* map_sigma, double* partial, unsigned* counters, double* per_sample) {
    __shared__ double red[3 * 8];
    __shared__ int flag;
    const int n = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
    const int Ho = H - k + 1, Wo = W - k + 1;
    const size_t off = (size_t)n * H * W;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long idx = (long long)blk * 256 + threadIdx.x;
    if (idx < (long long)Ho * Wo) {
        const int i0 = (int)(idx / Wo), j0 = (int)(idx % Wo);
        const GenMoments m = gen_moments(X + off, Y + off, W, i0, j0, k, taps);
        const double vx = fmax(m.exx - m.mx * m.mx, 0.0), vy = fmax(m.eyy - m.my * m.my, 0.0), cov = m.exy - m.mx * m.my;
        const double A1 = 2.0 * m.mx * m.my + C1, B1 = m.mx * m.mx + m.my * m.my + C1, A2 = 2.0 * cov + C2, B2 = vx + vy + C2;
        acc[0] = (A1 * A2) / (B1 * B2);
        acc[1] = A2 / B2;
        acc[2] = fmax(vx, 1e-4);
        const size_t o = (size_t)n * Ho * Wo + idx;
        if (map_ssim) map_ssim[o] = (float)acc[0];
        if (map_cs) map_cs[o] = (float)acc[1];
        if (map_sigma) map_sigma[o] = (float)acc[2];
    }
    if (!per_sample) return;
    double t[3];
    if (!block_finish<3, 256>(acc, red, &flag, partial, counters, n, blk, nblk, t)) return;
    const double inv = 1.0 / ((double)Ho * (double)Wo);
    per_sample[3 * n + 0] = t[0] * inv; per_sample[3 * n + 1] = t[1] * inv; per_sample[3 * n + 2] = t[2] * inv;
}

// backward, pass 1: d(objective)/d(moments) per window position -> coef[pos][5] = d/d(mu_x, mu_y, E[x^2], E[y^2], E[xy]),
// objective = sum_pos g_ssim ssim + g_cs cs + g_sigma sigma with the upstream weights per sample (x 1 / (Ho Wo)) or per position.
__global__ void __launch_bounds__(256)
ssim_generic_coef_kernel(const float* __restrict__ X, const float* __restrict__ Y, int H, int W, int k, const Taps taps, double C1, double C2,
                         const float* __restrict__ g_ssim, const float* __restrict__ g_cs, const float* __restrict__ g_sigma, int maps,
                         double* __restrict__ coef) {
    const int n = blockIdx.y;
    const int Ho = H - k + 1, Wo = W - k + 1;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)Ho * Wo) return;
    const size_t off = (size_t)n * H * W;
    const int i0 = (int)(idx / Wo), j0 = (int)(idx % Wo);
    const GenMoments m = gen_moments(X + off, Y + off, W, i0, j0, k, taps);
    const size_t o = (size_t)n * Ho * Wo + idx;
    const double scale = maps ? 1.0 : 1.0 / ((double)Ho * (double)Wo);
    const double gs = g_ssim ? scale * (double)g_ssim[maps ? o : n] : 0.0;
    const double gc = g_cs ? scale * (double)g_cs[maps ? o : n] : 0.0;
    const double gg = g_sigma ? scale * (double)g_sigma[maps ? o : n] : 0.0;
    const double vxr = m.exx - m.mx * m.mx, vyr = m.eyy - m.my * m.my;
    const double ix = vxr >= 0.0 ? 1.0 : 0.0, iy = vyr >= 0.0 ? 1.0 : 0.0;      // clamp(min=0) passes the gradient at the boundary
    const double vx = fmax(vxr, 0.0), vy = fmax(vyr, 0.0), cov = m.exy - m.mx * m.my;
    const double A1 = 2.0 * m.mx * m.my + C1, B1 = m.mx * m.mx + m.my * m.my + C1, A2 = 2.0 * cov + C2, B2 = vx + vy + C2;
    const double Lm = A1 / B1, CS = A2 / B2;
    const double dJ_dL = gs * CS, dJ_dCS = gs * Lm + gc;
    const double dL_dmx = (2.0 * m.my - Lm * 2.0 * m.mx) / B1, dL_dmy = (2.0 * m.mx - Lm * 2.0 * m.my) / B1;
    const double dCS_dcov = 2.0 / B2, dCS_dv = -CS / B2;
    const double isg = (vx >= 1e-4) ? 1.0 : 0.0;                                 // sigma = clamp(var_x, 1e-4)
    const double dJ_dvx = (dJ_dCS * dCS_dv + gg * isg) * ix, dJ_dvy = dJ_dCS * dCS_dv * iy, dJ_dcov = dJ_dCS * dCS_dcov;
    double* c = coef + o * 5;
    c[0] = dJ_dL * dL_dmx - dJ_dcov * m.my - 2.0 * m.mx * dJ_dvx;   // d/d mu_x
    c[1] = dJ_dL * dL_dmy - dJ_dcov * m.mx - 2.0 * m.my * dJ_dvy;   // d/d mu_y
    c[2] = dJ_dvx;                                                  // d/d E[x^2]
    c[3] = dJ_dvy;                                                  // d/d E[y^2]
    c[4] = dJ_dcov;                                                 // d/d E[xy]
}

// backward, pass 2: every pixel gathers the windows that cover it
__global__ void __launch_bounds__(256)
ssim_generic_gather_kernel(const float* __restrict__ X, const float* __restrict__ Y, int H, int W, int k, const Taps taps,
                           const double* __restrict__ coef, float* __restrict__ dX, float* __restrict__ dY) {
    const int n = blockIdx.y;
    const int Ho = H - k + 1, Wo = W - k + 1;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)H * W) return;
    const int i = (int)(idx / W), j = (int)(idx % W);
    const size_t p = (size_t)n * H * W + idx;
    const double x = X[p], y = Y[p];
    double gx = 0.0, gy = 0.0;
    for (int u = 0; u < k; ++u) {
        const int qi = i - u;
        if (qi < 0 || qi >= Ho) continue;
        for (int v = 0; v < k; ++v) {
            const int qj = j - v;
            if (qj < 0 || qj >= Wo) continue;
            const double w = (double)(float)(taps.w[u] * taps.w[v]);
            const double* c = coef + ((size_t)n * Ho * Wo + (size_t)qi * Wo + qj) * 5;
            gx += w * (c[0] + 2.0 * x * c[2] + y * c[4]);
            gy += w * (c[1] + 2.0 * y * c[3] + x * c[4]);
        }
    }
    if (dX) dX[p] = (float)gx;
    if (dY) dY[p] = (float)gy;
}

static int gen_check(const void* x, const void* y, int B, int H, int W, int win) {
    if (!x || !y) { set_error("null image pointer"); return MMIF_E_NULL; }
    if (win < 1 || win > kMaxWin) { set_error("SSIM window %d out of range 1..%d", win, kMaxWin); return MMIF_E_MODE; }
    if (B < 1 || H < win || W < win) { set_error("shape (%d,%d,%d) smaller than the %d-tap window", B, H, W, win); return MMIF_E_SHAPE; }
    if (((uintptr_t)x | (uintptr_t)y) & 3) { set_error("image pointers must be 4-byte aligned"); return MMIF_E_ALIGN; }
    return MMIF_OK;
}

}  // namespace mmif

using namespace mmif;

extern "C" size_t mmif_ssim_generic_workspace_bytes(int B, int H, int W, int win) {
    if (B < 1 || win < 1 || H < win || W < win) return 0;
    const long long npos = (long long)(H - win + 1) * (W - win + 1);
    const size_t nblk = (size_t)((npos + 255) / 256);
    return ws_counters_bytes(B) + (size_t)B * nblk * 3 * sizeof(double);
}
extern "C" size_t mmif_ssim_generic_coef_doubles(int B, int H, int W, int win) {
    if (B < 1 || win < 1 || H < win || W < win) return 0;
    return (size_t)B * (size_t)(H - win + 1) * (size_t)(W - win + 1) * 5;
}

extern "C" int mmif_ssim_generic_fwd(const float* x, const float* y, int B, int H, int W, int win, double sigma, float data_range,
                                     double* per_sample3, float* map_ssim, float* map_cs, float* map_sigma, void* ws, size_t ws_bytes,
                                     void* stream) {
    int rc = gen_check(x, y, B, H, W, win);
    if (rc) return rc;
    if (!per_sample3 && !map_ssim && !map_cs && !map_sigma) { set_error("no output requested"); return MMIF_E_NULL; }
    const size_t need = mmif_ssim_generic_workspace_bytes(B, H, W, win);
    if (per_sample3 && (!ws || ws_bytes < need)) { set_error("workspace too small: %zu < %zu", ws_bytes, need); return MMIF_E_WORKSPACE; }
    Taps taps;
    make_taps(&taps, win, sigma);
    const double R = data_range;
    const long long npos = (long long)(H - win + 1) * (W - win + 1);
    dim3 grid((unsigned)((npos + 255) / 256), B);
    ssim_generic_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, H, W, win, taps, (0.01 * R) * (0.01 * R), (0.03 * R) * (0.03 * R),
                                                                     map_ssim, map_cs, map_sigma,
                                                                     per_sample3 ? (double*)((unsigned char*)ws + ws_counters_bytes(B)) : nullptr,
                                                                     (unsigned*)ws, per_sample3);
    count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}

extern "C" int mmif_ssim_generic_bwd(const float* x, const float* y, int B, int H, int W, int win, double sigma, float data_range,
                                     const float* g_ssim, const float* g_cs, const float* g_sigma, int maps, float* dx, float* dy,
                                     double* coef, void* stream) {
    int rc = gen_check(x, y, B, H, W, win);
    if (rc) return rc;
    if (!coef || (!dx && !dy)) { set_error("null coef / no gradient requested"); return MMIF_E_NULL; }
    if (!g_ssim && !g_cs && !g_sigma) { set_error("no upstream gradient given"); return MMIF_E_NULL; }
    Taps taps;
    make_taps(&taps, win, sigma);
    const double R = data_range;
    const long long npos = (long long)(H - win + 1) * (W - win + 1);
    cudaStream_t st = (cudaStream_t)stream;
    ssim_generic_coef_kernel<<<dim3((unsigned)((npos + 255) / 256), B), 256, 0, st>>>(x, y, H, W, win, taps, (0.01 * R) * (0.01 * R),
                                                                                     (0.03 * R) * (0.03 * R), g_ssim, g_cs, g_sigma, maps, coef);
    count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    const long long npix = (long long)H * W;
    ssim_generic_gather_kernel<<<dim3((unsigned)((npix + 255) / 256), B), 256, 0, st>>>(x, y, H, W, win, taps, coef, dx, dy);
    count_launch(MMIF_CNT_AUX);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
