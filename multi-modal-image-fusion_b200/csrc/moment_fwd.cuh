// Forward strip-streaming kernel over the engine of stencil.cuh, templated on the window size and
// on the per-window epilogue:
//   EPI_SSIM : SSIM / CS / clamped-variance sums of the pairs (x1,y), (x2,y)  [+ optional Sobel and
//              pixel terms of the fusion objective, loss.py:294-344]
//   EPI_VIF  : VIF numerator / denominator sums of the pairs and of the gain-selected source
//              (metric.py:436-456, 486-489)
// Reads x1, x2, y once (12 B/pixel) and writes only per-CTA partial sums; the last CTA of each sample
// reduces that sample's partials in a fixed order, and (FIN_LOSS) the last sample-finisher writes the
// loss scalars — one launch, deterministic, no float atomics.
#pragma once
#include "stencil.cuh"

namespace mmif {

enum { EPI_SSIM = 0, EPI_VIF = 1, EPI_MSW = 2, EPI_MAPS = 3 };
enum { FIN_SUMS = 0, FIN_LOSS = 1 };

// TMA needs the global address of a box (innermost coordinate * 4 B) 16-byte aligned, so strips
// advance by a multiple of 4 columns.
template <int WIN>
struct FwdGeo {
    static constexpr int HALO = WIN - 1;
    static constexpr int TWO = ((kTWI - HALO) / 4) * 4;   // output columns per strip
};

// Everything the "last CTA finishes" logic needs; shared by the forward kernel and by the
// single-pass (loss + gradient) mode of the backward kernel.
struct FinishParams {
    int B, H, W, Hout, Wout;
    int finalize;            // FIN_SUMS / FIN_LOSS
    float w_ssim, w_pixel, w_grad;
    unsigned* counters;      // [B+1], zero on entry, left zero
    double* partial;         // [B][nblk][8]
    double* sums;            // per-sample raw sums: sums[n*sums_stride + 0..7]
    long long sums_stride;
    double* out;             // FIN_LOSS: loss block (mmif_b200.h MMIF_LOSS_*), followed by its float32 mirror
};
// float32 mirror of the loss block (same indices), stored right behind the doubles: what a float32 caller
// (the autograd wrapper) reads without a conversion kernel.
static inline size_t loss_block_doubles(int B) { return (size_t)MMIF_LOSS_HEAD + (size_t)(B > 0 ? B : 0) * MMIF_LOSS_PER_SAMPLE; }

#ifdef __CUDACC__
// v[] = this thread's 8 partial sums ([ssim1, ssim2, cs1, cs2, sig1, sig2, pix, grad] for the SSIM
// epilogue).  Block-reduce, publish, and let the last CTA of sample n re-reduce that sample in a fixed
// order; with FIN_LOSS the last sample-finisher writes the loss scalars.  All threads must call it.
template <int NTHREADS>
__device__ __forceinline__ void cta_finish(const FinishParams& p, double (&v)[8], double* red, int* flag, int n, int blk, int nblk) {
    block_sum<8, NTHREADS>(v, red);
    if (threadIdx.x == 0) {
        double* dst = p.partial + ((size_t)n * nblk + blk) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = v[i];
        __threadfence();
        const unsigned prev = atomicAdd(&p.counters[n], 1u);
        *flag = (prev == (unsigned)(nblk - 1));
    }
    __syncthreads();
    if (!*flag) return;
    __threadfence();
    double t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = 0.0;
    for (int k = threadIdx.x; k < nblk; k += NTHREADS) {
        const double* srcp = p.partial + ((size_t)n * nblk + k) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] += __ldcg(srcp + i);
    }
    block_sum<8, NTHREADS>(t, red);
    if (threadIdx.x == 0) {
        double* sums = p.sums + (size_t)n * p.sums_stride;
#pragma unroll
        for (int i = 0; i < 8; ++i) sums[i] = t[i];
        p.counters[n] = 0u;
        if (p.finalize == FIN_LOSS) {
            const double inv = 1.0 / ((double)p.Hout * (double)p.Wout);
            double* so = p.out + MMIF_LOSS_HEAD + (size_t)n * MMIF_LOSS_PER_SAMPLE;
            so[0] = t[0] * inv; so[1] = t[2] * inv; so[2] = t[4] * inv;   // ssim1, cs1, sigma1
            so[3] = t[1] * inv; so[4] = t[3] * inv; so[5] = t[5] * inv;   // ssim2, cs2, sigma2
            float* out32 = reinterpret_cast<float*>(p.out + (size_t)MMIF_LOSS_HEAD + (size_t)p.B * MMIF_LOSS_PER_SAMPLE);
            float* so32 = out32 + MMIF_LOSS_HEAD + (size_t)n * MMIF_LOSS_PER_SAMPLE;
#pragma unroll
            for (int i = 0; i < 6; ++i) so32[i] = (float)so[i];
            __threadfence();
            const unsigned prev = atomicAdd(&p.counters[p.B], 1u);
            if (prev == (unsigned)(p.B - 1)) {
                __threadfence();
                double a1 = 0.0, a2 = 0.0, px = 0.0, gr = 0.0;
                for (int k = 0; k < p.B; ++k) {
                    const double* q = p.sums + (size_t)k * p.sums_stride;
                    a1 += __ldcg(q + 0) * inv; a2 += __ldcg(q + 1) * inv; px += __ldcg(q + 6); gr += __ldcg(q + 7);
                }
                const double npx = (double)p.B * (double)p.H * (double)p.W;
                const double l_ssim = (double)p.w_ssim * (1.0 - 0.5 * (a1 / p.B + a2 / p.B));
                const double l_pix = (double)p.w_pixel * px / npx;
                const double l_grad = (double)p.w_grad * gr / npx;
                p.out[MMIF_LOSS_SSIM] = l_ssim;
                p.out[MMIF_LOSS_PIXEL] = l_pix;
                p.out[MMIF_LOSS_GRAD] = l_grad;
                p.out[MMIF_LOSS_TOTAL] = l_ssim + l_pix + l_grad;
                out32[MMIF_LOSS_SSIM] = (float)l_ssim;
                out32[MMIF_LOSS_PIXEL] = (float)l_pix;
                out32[MMIF_LOSS_GRAD] = (float)l_grad;
                out32[MMIF_LOSS_TOTAL] = (float)(l_ssim + l_pix + l_grad);
                p.counters[p.B] = 0u;
            }
        }
    }
}
#endif

struct FwdParams {
    const float* x1; const float* x2; const float* y;
    FinishParams fin;
    int B, H, W, Hout, Wout;
    int seg_rows, nseg, nstrip;
    Taps taps;
    float C1, C2;
    int pixel_combine, grad_combine, pixel_norm, grad_norm;
    float w_ssim, w_pixel, w_grad;
    int use_tma;
    int do_sobel;            // EPI_SSIM only: also accumulate the Sobel / pixel terms
    float* maps[6];          // EPI_MAPS: [ssim1, cs1, sigma1, ssim2, cs2, sigma2] maps [B][Hout][Wout] (each may be NULL)
    unsigned char* denorm;   // do_sobel: also store uint8(clip(y, 0, 1) * 255) [B][H][W] (test.py:70-73 post-step) or NULL
    int plain_moments;       // 1: exact central moments (window sum taken as 1: no eps / rho emulation of the reference's fp32 window)
};

static inline size_t ws_counters_bytes(int B) { return (size_t)(((B + 1) * 4 + 255) / 256) * 256; }

// --- forward Sobel / pixel terms for the pixel rows a batch owns (thread = column) ----------
using SmemF = SmemT<3, kTWI>;   // forward kernels: 3-slot ring, 72 KB per CTA -> 3 CTAs per SM

template <class SM>
__device__ __forceinline__ void fwd_sobel_pixel(const SM& sm, const FwdParams& p, int i0, int j0, int rlo, int rhi, int clo,
                                                int chi, float& pix_sum, float& grad_sum) {
    const int c = j0 + (int)threadIdx.x;
    if (c < clo || c >= chi) return;
    const int t0 = threadIdx.x;
    const int tm = ((c == 0) ? 1 : c - 1) - j0;
    const int tp = ((c == p.W - 1) ? p.W - 2 : c + 1) - j0;
    float dA[3], dB[3], sA[3], sB[3], u0p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) dA[k] = dB[k] = sA[k] = sB[k] = u0p[k] = 0.f;
    for (int r = rlo - 1; r <= rhi; ++r) {
        const int rr = (r < 0) ? -r : ((r >= p.H) ? 2 * p.H - 2 - r : r);
        const int lr = SM::wrap_any(rr - i0);
        float S[3], u0[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float um = sm.ring[k][lr][tm];
            const float uc = sm.ring[k][lr][t0];
            const float up = sm.ring[k][lr][tp];
            const float d = up - um;
            const float s = um + 2.f * uc + up;
            const float gx = dA[k] + 2.f * dB[k] + d;
            const float gy = s - sA[k];
            S[k] = fabsf(gx) + fabsf(gy);
            dA[k] = dB[k]; dB[k] = d; sA[k] = sB[k]; sB[k] = s;
            u0[k] = u0p[k]; u0p[k] = uc;        // u0 = centre value of row r-1
        }
        if (r >= rlo + 1) {                      // S[] / u0[] now describe pixel row r-1
            if (p.grad_combine == MMIF_COMBINE_MAX) {
                grad_sum += norm_val(S[2] - fmaxf(S[0], S[1]), p.grad_norm);
            } else {
                grad_sum += 0.5f * (norm_val(S[2] - S[0], p.grad_norm) + norm_val(S[2] - S[1], p.grad_norm));
            }
            if (p.pixel_combine == MMIF_COMBINE_MAX) {
                pix_sum += norm_val(u0[2] - fmaxf(u0[0], u0[1]), p.pixel_norm);
            } else {
                pix_sum += 0.5f * (norm_val(u0[2] - u0[0], p.pixel_norm) + norm_val(u0[2] - u0[1], p.pixel_norm));
            }
            if (p.denorm) {      // denorm() of data/transform.py:32-35: clip(0,1) * 255 in float32, truncated to uint8 (NaN -> 0)
                const float v = fminf(fmaxf(u0[2], 0.f), 1.f) * 255.0f;
                p.denorm[((size_t)blockIdx.z * p.H + (r - 1)) * p.W + c] = (unsigned char)__float2uint_rz(v);
            }
        }
    }
}

// Same terms for the training configuration (mode='max', 'l1': train.py:67-68,307-308) without the
// run-time mode switches; the two sources are processed as one packed (x1, x2) pair.
template <class SM>
__device__ __forceinline__ void fwd_sobel_pixel_fast(const SM& sm, int H, int W, int i0, int j0, int rlo, int rhi, int clo,
                                                     int chi, float& pix_sum, float& grad_sum) {
    const int c = j0 + (int)threadIdx.x;
    if (c < clo || c >= chi) return;
    const int t0 = threadIdx.x;
    const int tm = ((c == 0) ? 1 : c - 1) - j0;
    const int tp = ((c == W - 1) ? W - 2 : c + 1) - j0;
    float2 dA = f2(0.f, 0.f), dB = dA, sA = dA, sB = dA, ucp = dA;
    float dAy = 0.f, dBy = 0.f, sAy = 0.f, sBy = 0.f, ucpy = 0.f;
    const float2 two = bcast(2.f), neg1 = bcast(-1.f);
    for (int r = rlo - 1; r <= rhi; ++r) {
        const int rr = (r < 0) ? -r : ((r >= H) ? 2 * H - 2 - r : r);
        const int lr = SM::wrap_any(rr - i0);
        const float2 um = f2(sm.ring[0][lr][tm], sm.ring[1][lr][tm]);
        const float2 uc = f2(sm.ring[0][lr][t0], sm.ring[1][lr][t0]);
        const float2 up = f2(sm.ring[0][lr][tp], sm.ring[1][lr][tp]);
        const float umy = sm.ring[2][lr][tm], ucy = sm.ring[2][lr][t0], upy = sm.ring[2][lr][tp];
        const float2 d = fma2(neg1, um, up);
        const float2 s = fma2(two, uc, add2(um, up));
        const float2 gx = fma2(two, dB, add2(dA, d));
        const float2 gy = fma2(neg1, sA, s);
        const float dy = upy - umy;
        const float sy = fmaf(2.f, ucy, umy + upy);
        const float gxy = fmaf(2.f, dBy, dAy + dy);
        const float gyy = sy - sAy;
        if (r >= rlo + 1) {                      // the sums describe pixel row r-1
            const float S1 = fabsf(gx.x) + fabsf(gy.x), S2 = fabsf(gx.y) + fabsf(gy.y), Sy = fabsf(gxy) + fabsf(gyy);
            grad_sum += fabsf(Sy - fmaxf(S1, S2));
            pix_sum += fabsf(ucpy - fmaxf(ucp.x, ucp.y));
        }
        dA = dB; dB = d; sA = sB; sB = s; ucp = uc;
        dAy = dBy; dBy = dy; sAy = sBy; sBy = sy; ucpy = ucy;
    }
}

// VIF sums of log2(1 + x) (metric.py:454-456) are kept as running PRODUCTS of (1 + x) in double
// (mantissa in [1,2) + an integer exponent, renormalised once per batch of 8 factors): the FP64 pipe is
// idle in this kernel, the FP32 pipe is its bottleneck, and a logarithm is taken once per thread at
// the end instead of four times per pixel.  Exact to double rounding, i.e. nearer the reference's
// fp64 evaluation than its fp32 log2(1 + x), which loses the low bits of a small x.
struct LogProd {
    double m; int e;
    __device__ __forceinline__ void init() { m = 1.0; e = 0; }
    __device__ __forceinline__ void mul(double f) { m *= f; }
    __device__ __forceinline__ void renorm() {          // m >= 1 always (factors are >= 1)
        int hi = __double2hiint(m);
        const int lo = __double2loint(m);
        const int ex = (hi >> 20) - 1023;
        if (ex < 1024) { e += ex; hi -= ex << 20; m = __hiloint2double(hi, lo); }     // inf / nan stay as they are
    }
    __device__ __forceinline__ double log2_total() const { return (double)e + log2(m); }
};

template <int WIN, int EPI>
__global__ void __launch_bounds__(kNT, 3)
moment_fwd_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                  const __grid_constant__ CUtensorMap mapy, const FwdParams p) {
    constexpr int HALO = FwdGeo<WIN>::HALO;
    constexpr int TWO = FwdGeo<WIN>::TWO;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
    const int strip = blockIdx.x, seg = blockIdx.y, n = blockIdx.z;
    const int j0 = strip * TWO, i0 = seg * p.seg_rows;
    const int rows_out = min(p.seg_rows, p.Hout - i0);
    const int cols_out = min(TWO, p.Wout - j0);
    const int nb = (rows_out + kRB - 1) / kRB;
    const size_t img_off = (size_t)n * p.H * p.W;

    RingSrc src;
    src.img[0] = p.x1 + img_off; src.img[1] = p.x2 + img_off; src.img[2] = p.y + img_off;
    src.H = p.H; src.W = p.W; src.row0 = i0; src.col0 = j0; src.n = n; src.use_tma = p.use_tma != 0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) mbar_init((uint64_t*)&sm.mbar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    Shift sh = tile_shift(sm, src.img[0], src.img[1], src.img[2], p.H, p.W, i0, rows_out + HALO, j0, p.taps);
    if (p.plain_moments) sh = make_shift(sh.c.x, sh.c.y, sh.cy, 1.f, 0.f, 0.f);

    ring_issue(sm, src, &map1, &map2, &mapy, 0);
    ring_issue(sm, src, &map1, &map2, &mapy, 1);
    ring_issue(sm, src, &map1, &map2, &mapy, 2);
    __syncthreads();
    ring_wait(sm, src, 0);
    ring_wait(sm, src, 1);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ho = lane & 7, hg = warp * 4 + (lane >> 3);
    float2 s0 = f2(0.f, 0.f), s1 = f2(0.f, 0.f), s2 = f2(0.f, 0.f);   // three packed running sums
    float pix_sum = 0.f, grad_sum = 0.f;
    LogProd lp[6];                                                     // EPI_VIF: num1, den1, num2, den2, numsel, densel
#pragma unroll
    for (int i = 0; i < 6; ++i) lp[i].init();
    const int clo = (strip == 0) ? 0 : j0 + HALO / 2;
    const int chi = (strip == p.nstrip - 1) ? p.W : j0 + TWO + HALO / 2;
    constexpr float kVifEps = 1e-10f, kVifNoise = 325.125f;            // metric.py:407-408

    unsigned cmask = 0u;                         // bit j: window column hg*8 + j exists
#pragma unroll
    for (int j = 0; j < 8; ++j) cmask |= (hg * 8 + j < cols_out) ? (1u << j) : 0u;
    const bool fast_terms = p.pixel_combine == MMIF_COMBINE_MAX && p.grad_combine == MMIF_COMBINE_MAX &&
                            p.pixel_norm == MMIF_NORM_L1 && p.grad_norm == MMIF_NORM_L1 && p.denorm == nullptr;
    for (int b = 0; b < nb; ++b) {
        ring_wait(sm, src, b + 2);
        vpass_moments<WIN>(sm, p.taps, sh, (b % 3) * kRB);
        if (EPI == EPI_SSIM && p.do_sobel) {
            const int rlo = (seg == 0 && b == 0) ? 0 : i0 + b * kRB + HALO / 2;
            const int rhi = (seg == p.nseg - 1 && b == nb - 1) ? p.H : i0 + b * kRB + kRB + HALO / 2;
            if (fast_terms) fwd_sobel_pixel_fast(sm, p.H, p.W, i0, j0, rlo, rhi, clo, chi, pix_sum, grad_sum);
            else fwd_sobel_pixel(sm, p, i0, j0, rlo, rhi, clo, chi, pix_sum, grad_sum);
        }
        __syncthreads();
        // group b is dead now (V-pass and Sobel rows of batch b+1 start at group b+1): refill its slot
        if (b + 1 < nb) ring_issue(sm, src, &map1, &map2, &mapy, b + 3);
        const int rows_b = min(kRB, rows_out - b * kRB);
        if (ho < rows_b && hg * 8 < cols_out) {
            float2 acc[8][4];
            hpass<WIN, 4, false>(sm.vbuf + ho * kVPitch + hg * 8, kVCols, p.taps, acc);
            // Branch-free over the 8 columns (their chains interleave); columns past the last window are
            // computed on zero-filled data and dropped by a select.
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool vj = (cmask >> j) & 1u;
                const Stats st = stats_from(moments_of(acc[j]), sh);
                const float2 vk = max2(st.vk, 0.f);
                const float vy = fmaxf(st.vy, 0.f);
                if (EPI == EPI_VIF) {
                    double fn[2], fd[2];
                    float gg[2];
                    const float v1[2] = {vk.x, vk.y}, c12[2] = {st.cov.x, st.cov.y};
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        float sig1 = v1[k];
                        float g = fdiv_nr(c12[k], sig1 + kVifEps);
                        float sv = vy - g * c12[k];
                        if (sig1 < kVifEps) { g = 0.f; sv = vy; sig1 = 0.f; }
                        if (vy < kVifEps) { g = 0.f; sv = 0.f; }
                        if (g < 0.f) { sv = vy; g = 0.f; }
                        if (sv < kVifEps) sv = kVifEps;
                        fn[k] = 1.0 + (double)fdiv_nr(g * g * sig1, sv + kVifNoise);
                        fd[k] = 1.0 + (double)(sig1 * (1.0f / kVifNoise));
                        gg[k] = g;
                    }
                    const bool pick1 = gg[0] < gg[1];
                    if (vj) {        // predicated multiplies (a column past the last window contributes the factor 1)
                        lp[0].mul(fn[0]); lp[1].mul(fd[0]); lp[2].mul(fn[1]); lp[3].mul(fd[1]);
                        lp[4].mul(pick1 ? fn[0] : fn[1]); lp[5].mul(pick1 ? fd[0] : fd[1]);
                    }
                } else {
                    const float2 A1 = fma2(st.mu, bcast(2.f * st.muy), bcast(p.C1));
                    const float2 B1 = fma2(st.mu, st.mu, bcast(fmaf(st.muy, st.muy, p.C1)));
                    const float2 A2 = fma2(bcast(2.f), st.cov, bcast(p.C2));
                    const float2 B2 = add2(vk, bcast(vy + p.C2));
                    const float2 CS = mul2(A2, rcp2(B2));
                    const float2 S = mul2(mul2(A1, rcp2(B1)), CS);
                    const float2 SG = max2(vk, 1e-4f);
                    if (EPI == EPI_MSW) {               // gamma-weighted SSIM of the two pairs (loss.py:230-235)
                        const float gm = __fdiv_rn(SG.x, fmaxf(SG.x + SG.y, 1e-7f));
                        s0.x += vj ? (gm * S.x + (1.f - gm) * S.y) : 0.f;
                    } else if (EPI == EPI_MAPS) {       // size_average=False: the maps themselves (loss.py:99-108)
                        if (vj) {
                            const size_t o = ((size_t)n * p.Hout + (i0 + b * kRB + ho)) * p.Wout + (j0 + hg * 8 + j);
                            if (p.maps[0]) p.maps[0][o] = S.x;
                            if (p.maps[1]) p.maps[1][o] = CS.x;
                            if (p.maps[2]) p.maps[2][o] = SG.x;
                            if (p.maps[3]) p.maps[3][o] = S.y;
                            if (p.maps[4]) p.maps[4][o] = CS.y;
                            if (p.maps[5]) p.maps[5][o] = SG.y;
                        }
                    } else if (vj) {          // predicated adds instead of six selects per column
                        s0 = add2(s0, S);
                        s1 = add2(s1, CS);
                        s2 = add2(s2, SG);
                    }
                }
            }
            if (EPI == EPI_VIF) {
#pragma unroll
                for (int i = 0; i < 6; ++i) lp[i].renorm();
            }
        }
        __syncthreads();
    }

    // ---- per-CTA partial, then deterministic per-sample / global finish by the last arrivals ----
    // SSIM: [ssim1, ssim2, cs1, cs2, sig1, sig2, pix, grad]; VIF: [num1, den1, num2, den2, numsel, densel, 0, 0]
    double v[8] = {(double)s0.x, (double)s0.y, (double)s1.x, (double)s1.y, (double)s2.x, (double)s2.y,
                   (double)pix_sum, (double)grad_sum};
    if (EPI == EPI_VIF) {
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = lp[i].log2_total();
    }
    cta_finish<kNT>(p.fin, v, sm.red, &sm.flag, n, seg * p.nstrip + strip, p.nstrip * p.nseg);
}

// ---- host-side geometry + launcher (defined in moment_fwd.cu) ----------------------------------
struct FwdLaunch {
    int win;                 // 3, 5, 9, 11 or 17
    double sigma;
    int epi;                 // EPI_*
    int finalize;            // FIN_*
    int do_sobel;
    float data_range;
    MmifLossCfg cfg;         // combine / norm / weights (EPI_SSIM + do_sobel)
    float* maps[6];          // EPI_MAPS outputs
    unsigned char* denorm;   // do_sobel: uint8 image of y (or NULL)
    int plain_moments;       // see FwdParams
};
int pick_seg_rows(int rows, int other_ctas, int slots, int extra_rows, double tail = 0.0);
int fwd_seg_rows(int rows, int other_ctas);
size_t fwd_ws_bytes(int win, int B, int H, int W);       // counters + partials (sums live elsewhere)
// ws: zero-initialised workspace of fwd_ws_bytes; sums: B x sums_stride doubles (device).
int launch_moment_fwd(const FwdLaunch& L, const float* x1, const float* x2, const float* y, int B, int H, int W,
                      double* sums, long long sums_stride, double* out, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace mmif
