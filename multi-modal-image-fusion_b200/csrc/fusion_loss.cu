// Fusion objective (reference core/loss.py): SSIM('ssim') + PixelLoss + GradLoss, forward and
// backward, as two strip-streaming kernels (see stencil.cuh for the engine).
//
//   fusion_loss_fwd : reads I1, I2, If once (12 B/pixel), writes per-CTA partial sums only; the last
//                     CTA of each sample reduces that sample in a fixed order (deterministic), the
//                     last sample-finisher writes the loss scalars.  One launch.
//   fusion_loss_bwd : recomputes the moments, forms the per-window derivative coefficients, runs the
//                     adjoint blur and the folded Sobel adjoint and writes dL/dIf once
//                     (12 B read + 4 B written per pixel).  One launch.
#include "moment_fwd.cuh"

namespace mmif {

constexpr int WIN = 11;
constexpr int HALO = WIN - 1;
constexpr int kTWO = FwdGeo<WIN>::TWO;    // 116 SSIM-map columns per forward strip
constexpr int kOFF = 12;                  // backward tile origin = j0 - 12 (TMA: multiple of 4 columns, >= HALO)
constexpr int kTG = 104;                  // gradient columns per backward strip (104 + 12 + 10 <= 128, multiple of 4)
constexpr int kWC = kTWI - HALO;          // 118 window columns carry valid moments in a backward tile

struct BwdGeom { int Hout, Wout, seg_rows, nseg, nstrip; };
static BwdGeom bwd_geom(int B, int H, int W) {
    BwdGeom g;
    g.Hout = H - HALO; g.Wout = W - HALO;
    g.nstrip = ceil_div(W, kTG);
    g.seg_rows = fwd_seg_rows(H, B * g.nstrip);
    g.nseg = ceil_div(H, g.seg_rows);
    return g;
}

// =============================================================================== backward
struct BwdParams {
    const float* x1; const float* x2; const float* y;
    float* dF;
    const float* gout;       // 3 upstream gradients (device)
    int B, H, W, Hout, Wout;
    int seg_rows, nseg, nstrip;
    Taps taps;
    float C1, C2;
    int pixel_combine, grad_combine, pixel_norm, grad_norm;
    float w_ssim, w_pixel, w_grad;
    int use_tma;
    int vec_store;
};

constexpr int kCPitch = 2 * kTWI + 2;   // float2 units per coefficient row (2 pair-maps x 128 + 16 B pad)
constexpr int kTMC = 112;               // tmaps / gbuf columns

struct SmemBwd {
    Smem s;                              // ring + vbuf (vbuf also hosts tmaps and the B1->B2 buffer)
    alignas(16) float2 cbuf[kRB * kCPitch];          // coefficient rows of the current batch: (a,b) and (c1,c2)
    float gbuf[kRB][kTMC];               // pixel + Sobel gradient of the batch rows
};

__global__ void __launch_bounds__(kNT, 2)
fusion_loss_bwd_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                       const __grid_constant__ CUtensorMap mapy, const BwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemBwd& sb = *reinterpret_cast<SmemBwd*>(smem_raw);
    Smem& sm = sb.s;
    const int strip = blockIdx.x, seg = blockIdx.y, n = blockIdx.z;
    const int j0 = strip * kTG, i0 = seg * p.seg_rows;
    const int jw0 = j0 - kOFF;             // first window / input column of the tile (multiple of 4: TMA)
    const int R0 = i0 - HALO;              // first window / input row of the segment
    const int iend = min(i0 + p.seg_rows, p.H);
    const int jend = min(j0 + kTG, p.W);
    const int nb = (iend - R0 + kRB - 1) / kRB;   // batch b emits gradient rows [R0+8b, R0+8b+8)
    const size_t img_off = (size_t)n * p.H * p.W;
    const int t = threadIdx.x;

    RingSrc src;
    src.img[0] = p.x1 + img_off; src.img[1] = p.x2 + img_off; src.img[2] = p.y + img_off;
    src.H = p.H; src.W = p.W; src.row0 = R0; src.col0 = jw0; src.n = n; src.use_tma = p.use_tma != 0;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) mbar_init((uint64_t*)&sm.mbar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const Shift sh = tile_shift(sm, src.img[0], src.img[1], src.img[2], p.H, p.W, R0, iend - R0 + HALO, jw0, p.taps);
    const float g_ssim = __ldg(p.gout + 0), g_pix = __ldg(p.gout + 1), g_grad = __ldg(p.gout + 2);
    const float npx = (float)p.B * (float)p.H * (float)p.W;
    const float k_ssim = g_ssim * p.w_ssim * (-0.5f) / ((float)p.B * (float)p.Hout * (float)p.Wout);
    const float k_pix = g_pix * p.w_pixel / npx * (p.pixel_combine == MMIF_COMBINE_MAX ? 1.f : 0.5f);
    const float k_grad = g_grad * p.w_grad / npx * (p.grad_combine == MMIF_COMBINE_MAX ? 1.f : 0.5f);

    ring_issue(sm, src, &map1, &map2, &mapy, 0);
    ring_issue(sm, src, &map1, &map2, &mapy, 1);
    ring_issue(sm, src, &map1, &map2, &mapy, 2);
    __syncthreads();
    ring_wait(sm, src, 0);
    ring_wait(sm, src, 1);

    const int lane = t & 31, warp = t >> 5;
    const int ho = lane & 7, hg = warp * 4 + (lane >> 3);
    float2* tmaps = sm.vbuf;                          // [10][kTMC] (tx,ty), alive between S1 and S2
    float2* tbuf = sm.vbuf;                           // [8][2][128]+pad, alive between B1 and B2
    float2 carry[HALO][2];                            // vertical adjoint state: pending gradient rows
#pragma unroll
    for (int d = 0; d < HALO; ++d) carry[d][0] = carry[d][1] = f2(0.f, 0.f);

    for (int b = 0; b < nb; ++b) {
        const int Rb = R0 + b * kRB;                  // first gradient row of this batch
        ring_wait(sm, src, b + 2);
        const bool emit = (Rb + kRB > i0);            // any owned gradient row in this batch
        // ---------------- S1: tx, ty on rows [Rb-1, Rb+9) x cols [j0-1, j0+109) ----------------
        if (emit) {
            if (t < kTG + 2) {
                const int c = j0 - 1 + t;
                const bool cvalid = (c >= 0) && (c < p.W);
                const int cc = min(max(c, 0), p.W - 1);
                const int t0 = cc - jw0;
                const int tmc = ((cc == 0) ? 1 : cc - 1) - jw0;
                const int tpc = ((cc == p.W - 1) ? p.W - 2 : cc + 1) - jw0;
                float dA[3], dB[3], sA[3], sB[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) dA[k] = dB[k] = sA[k] = sB[k] = 0.f;
                const int lr_lo = b * kRB - 2, lr_hi = b * kRB + 17;   // ring-local rows present
                for (int q = Rb - 2; q <= Rb + 9; ++q) {
                    int rr = (q < 0) ? -q : ((q >= p.H) ? 2 * p.H - 2 - q : q);
                    int lr = min(max(rr - R0, lr_lo), lr_hi) & (kRingRows - 1);
                    float gx[3], gy[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float um = sm.ring[k][lr][tmc];
                        const float uc = sm.ring[k][lr][t0];
                        const float up = sm.ring[k][lr][tpc];
                        const float d = up - um;
                        const float s = um + 2.f * uc + up;
                        gx[k] = dA[k] + 2.f * dB[k] + d;
                        gy[k] = s - sA[k];
                        dA[k] = dB[k]; dB[k] = d; sA[k] = sB[k]; sB[k] = s;
                    }
                    if (q >= Rb) {                     // gx/gy describe row q-1 in [Rb-1, Rb+9)
                        const int qr = q - 1;
                        float2 txy = f2(0.f, 0.f);
                        if (cvalid && qr >= 0 && qr < p.H) {
                            const float S1 = fabsf(gx[0]) + fabsf(gy[0]);
                            const float S2 = fabsf(gx[1]) + fabsf(gy[1]);
                            const float Sy = fabsf(gx[2]) + fabsf(gy[2]);
                            float r;
                            if (p.grad_combine == MMIF_COMBINE_MAX) r = k_grad * norm_der(Sy - fmaxf(S1, S2), p.grad_norm);
                            else r = k_grad * (norm_der(Sy - S1, p.grad_norm) + norm_der(Sy - S2, p.grad_norm));
                            txy = f2(r * sgn(gx[2]), r * sgn(gy[2]));
                        }
                        tmaps[(qr - (Rb - 1)) * kTMC + t] = txy;
                    }
                }
            }
        }
        __syncthreads();
        // ---------------- S2: folded Sobel adjoint + pixel term -> gbuf ------------------------
        if (emit && t < kTG) {
            const int j = j0 + t;
            auto tmget = [&](int qi, int qj) -> float2 {   // bounds-checked (border folds only)
                const int ri = qi - (Rb - 1), ci = qj - (j0 - 1);
                if (ri < 0 || ri >= 10 || ci < 0 || ci >= kTG + 2) return f2(0.f, 0.f);
                return tmaps[ri * kTMC + ci];
            };
            auto Gslow = [&](int pi, int pj) -> float {
                float g = 0.f;
#pragma unroll
                for (int dr = -1; dr <= 1; ++dr) {
                    const float sv = (dr == 0) ? 2.f : 1.f;
                    g += sv * (tmget(pi + dr, pj - 1).x - tmget(pi + dr, pj + 1).x);
                }
#pragma unroll
                for (int dc = -1; dc <= 1; ++dc) {
                    const float shh = (dc == 0) ? 2.f : 1.f;
                    g += shh * (tmget(pi - 1, pj + dc).y - tmget(pi + 1, pj + dc).y);
                }
                return g;
            };
            for (int o = 0; o < kRB; ++o) {
                const int i = Rb + o;
                float g = 0.f;
                if (i >= i0 && i < iend && j < jend) {
                    const float2* r0 = tmaps + (o + 0) * kTMC + t;     // row i-1, col j-1
                    const float2* r1 = tmaps + (o + 1) * kTMC + t;     // row i
                    const float2* r2 = tmaps + (o + 2) * kTMC + t;     // row i+1
                    const float2 a0 = r0[0], a1 = r0[1], a2 = r0[2];
                    const float2 b0 = r1[0], b2 = r1[2];
                    const float2 c0 = r2[0], c1 = r2[1], c2 = r2[2];
                    g = (a0.x - a2.x) + 2.f * (b0.x - b2.x) + (c0.x - c2.x)
                      + (a0.y + 2.f * a1.y + a2.y) - (c0.y + 2.f * c1.y + c2.y);
                    const bool ftop = (i == 1), fbot = (i == p.H - 2), fl = (j == 1), fr = (j == p.W - 2);
                    if (ftop | fbot | fl | fr) {
                        if (ftop) g += Gslow(-1, j);
                        if (fbot) g += Gslow(p.H, j);
                        if (fl) g += Gslow(i, -1);
                        if (fr) g += Gslow(i, p.W);
                        if (ftop && fl) g += Gslow(-1, -1);
                        if (ftop && fr) g += Gslow(-1, p.W);
                        if (fbot && fl) g += Gslow(p.H, -1);
                        if (fbot && fr) g += Gslow(p.H, p.W);
                    }
                    const int lr = (b * kRB + o) & (kRingRows - 1);
                    const float u1 = sm.ring[0][lr][t + kOFF], u2 = sm.ring[1][lr][t + kOFF], uy = sm.ring[2][lr][t + kOFF];
                    if (p.pixel_combine == MMIF_COMBINE_MAX) g += k_pix * norm_der(uy - fmaxf(u1, u2), p.pixel_norm);
                    else g += k_pix * (norm_der(uy - u1, p.pixel_norm) + norm_der(uy - u2, p.pixel_norm));
                }
                sb.gbuf[o][t] = g;
            }
        }
        __syncthreads();
        if (b + 1 < nb) ring_issue(sm, src, &map1, &map2, &mapy, b + 3);
        // ---------------- A1: vertical moments of window rows [Rb, Rb+8) ------------------------
        vpass_moments<WIN>(sm, p.taps, sh, (b & 3) * kRB);
        __syncthreads();
        // ---------------- A2: horizontal moments -> derivative coefficients -> cbuf -------------
        {
            const int q = Rb + ho;                     // window row
            float2 ab[8], cc[8];
            const bool active = (hg * 8 < kWC) && (q >= 0) && (q < p.Hout);
            if (active) {
                float2 acc[8][4];
                hpass<WIN, 4, false>(sm.vbuf + ho * kVPitch + hg * 8, kVCols, p.taps, acc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pw = hg * 8 + j;
                    const int pc = jw0 + pw;
                    ab[j] = cc[j] = f2(0.f, 0.f);
                    if (pw < kWC && pc >= 0 && pc < p.Wout) {
                        const Moments mo = moments_of(acc[j]);
                        const Stats st = stats_from(mo, sh);
                        const float myk = (st.vy >= 0.f) ? 1.f : 0.f;
                        const float2 vk = max2(st.vk, 0.f);
                        const float vy = fmaxf(st.vy, 0.f);
                        const float2 A1 = fma2(muls(2.f, st.mu), bcast(st.muy), bcast(p.C1));
                        const float2 B1 = fma2(st.mu, st.mu, bcast(fmaf(st.muy, st.muy, p.C1)));
                        const float2 A2 = fma2(bcast(2.f), st.cov, bcast(p.C2));
                        const float2 B2 = add2(vk, bcast(vy + p.C2));
                        const float2 rB1 = fdiv_nr2(bcast(1.f), B1);
                        const float2 rB2 = fdiv_nr2(bcast(1.f), B2);
                        const float2 rBB = mul2(rB1, rB2);
                        const float2 S = mul2(mul2(A1, A2), rBB);
                        const float2 dcov = mul2(muls(2.f, A1), rBB);                    // dS/dcov
                        const float2 dvar = muls(-myk, mul2(S, rB2));                    // dS/dvar_y
                        // dS/dmu_y (luminance path) = 2 mu_k A2/(B1 B2) - 2 mu_y S / B1
                        const float2 dmu = fma2(mul2(muls(2.f, st.mu), A2), rBB, muls(-2.f * st.muy, mul2(S, rB1)));
                        // a' = dmu - dvar (2 my' + 2 eps cy) - dcov (mk' + eps ck)
                        const float2 a = fma2(muls(-1.f, dcov), add2(mo.mk, sh.ec),
                                              fma2(muls(-1.f, dvar), bcast(2.f * mo.my + sh.k1y), dmu));
                        ab[j] = f2(a.x + a.y, dvar.x + dvar.y);
                        cc[j] = dcov;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) ab[j] = cc[j] = f2(0.f, 0.f);
            }
            float4* d0 = reinterpret_cast<float4*>(sb.cbuf + ho * kCPitch + hg * 8);
            float4* d1 = reinterpret_cast<float4*>(sb.cbuf + ho * kCPitch + kTWI + hg * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                d0[j] = make_float4(ab[2 * j].x, ab[2 * j].y, ab[2 * j + 1].x, ab[2 * j + 1].y);
                d1[j] = make_float4(cc[2 * j].x, cc[2 * j].y, cc[2 * j + 1].x, cc[2 * j + 1].y);
            }
        }
        __syncthreads();
        // ---------------- B1: vertical adjoint (stateful), thread = window column ---------------
        {
            float2 P[kRB + HALO][2];
#pragma unroll
            for (int d = 0; d < kRB + HALO; ++d) {
                P[d][0] = (d < HALO) ? carry[d][0] : f2(0.f, 0.f);
                P[d][1] = (d < HALO) ? carry[d][1] : f2(0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < kRB; ++r) {
                const float2 v0 = sb.cbuf[r * kCPitch + t];
                const float2 v1 = sb.cbuf[r * kCPitch + kTWI + t];
#pragma unroll
                for (int d = 0; d < WIN; ++d) {          // window row Rb+r feeds gradient rows Rb+r+d with w[d]
                    P[r + d][0] = fmas(p.taps.w[d], v0, P[r + d][0]);
                    P[r + d][1] = fmas(p.taps.w[d], v1, P[r + d][1]);
                }
            }
            if (emit) {
#pragma unroll
                for (int o = 0; o < kRB; ++o) {
                    tbuf[o * kCPitch + t] = P[o][0];
                    tbuf[o * kCPitch + kTWI + t] = P[o][1];
                }
            }
#pragma unroll
            for (int d = 0; d < HALO; ++d) { carry[d][0] = P[d + kRB][0]; carry[d][1] = P[d + kRB][1]; }
        }
        __syncthreads();
        // ---------------- B2: horizontal adjoint + combine + store ------------------------------
        if (emit) {
            const int i = Rb + ho;
            if (i >= i0 && i < iend && hg * 8 < kTG && j0 + hg * 8 < jend) {
                float2 acc[8][2];
                // gradient column g sums window columns [g + kOFF - HALO, g + kOFF] of the tile
                hpass<WIN, 2, true>(tbuf + ho * kCPitch + hg * 8 + (kOFF - HALO), kTWI, p.taps, acc);
                const int lr = (b * kRB + ho) & (kRingRows - 1);
                float outv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int tc = kOFF + hg * 8 + j;
                    const float x1s = sm.ring[0][lr][tc] - sh.c.x;
                    const float x2s = sm.ring[1][lr][tc] - sh.c.y;
                    const float ys = sm.ring[2][lr][tc] - sh.cy;
                    const float dS = acc[j][0].x + 2.f * ys * acc[j][0].y + x1s * acc[j][1].x + x2s * acc[j][1].y;
                    outv[j] = fmaf(k_ssim, dS, sb.gbuf[ho][hg * 8 + j]);
                }
                float* dst = p.dF + img_off + (size_t)i * p.W + j0 + hg * 8;
                if (p.vec_store && j0 + hg * 8 + 8 <= jend) {
                    reinterpret_cast<float4*>(dst)[0] = make_float4(outv[0], outv[1], outv[2], outv[3]);
                    reinterpret_cast<float4*>(dst)[1] = make_float4(outv[4], outv[5], outv[6], outv[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j0 + hg * 8 + j < jend) dst[j] = outv[j];
                }
            }
        }
        __syncthreads();
    }
}

// =============================================================================== host side
static int check_common(const void* a, const void* b, const void* c, int B, int H, int W) {
    if (!a || !b || !c) { set_error("null image pointer"); return MMIF_E_NULL; }
    if (B < 1 || H < WIN || W < WIN) { set_error("shape (%d,%d,%d): need B>=1 and H,W >= %d", B, H, W, WIN); return MMIF_E_SHAPE; }
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 3) { set_error("image pointers must be 4-byte aligned"); return MMIF_E_ALIGN; }
    return MMIF_OK;
}
static int check_cfg(const MmifLossCfg* cfg) {
    if (!cfg) { set_error("null cfg"); return MMIF_E_NULL; }
    auto okc = [](int v) { return v == MMIF_COMBINE_MAX || v == MMIF_COMBINE_AVG; };
    auto okn = [](int v) { return v == MMIF_NORM_L1 || v == MMIF_NORM_L2; };
    if (!okc(cfg->pixel_combine) || !okc(cfg->grad_combine) || !okn(cfg->pixel_norm) || !okn(cfg->grad_norm)) {
        set_error("unsupported combine/norm mode in MmifLossCfg");
        return MMIF_E_MODE;
    }
    return MMIF_OK;
}

}  // namespace mmif

using namespace mmif;

// workspace = [fwd counters + partials][B x 8 per-sample sums]
extern "C" size_t mmif_loss_workspace_bytes(int B, int H, int W) {
    const size_t f = fwd_ws_bytes(WIN, B, H, W);
    return f ? f + (size_t)B * 8 * sizeof(double) : 0;
}
extern "C" size_t mmif_loss_out_doubles(int B) { return (size_t)MMIF_LOSS_HEAD + (size_t)(B > 0 ? B : 0) * MMIF_LOSS_PER_SAMPLE; }

extern "C" int mmif_fusion_loss_fwd(const float* i1, const float* i2, const float* f, int B, int H, int W,
                                    const MmifLossCfg* cfg, double* out, float* dF_unit, void* ws, size_t ws_bytes,
                                    void* stream) {
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    rc = check_cfg(cfg);
    if (rc) return rc;
    if (!out) { set_error("null out"); return MMIF_E_NULL; }
    if (cfg->want_grad || dF_unit) { set_error("single-pass gradient (want_grad) is not built yet"); return MMIF_E_MODE; }
    const size_t fws = fwd_ws_bytes(WIN, B, H, W);
    if (!ws || ws_bytes < mmif_loss_workspace_bytes(B, H, W)) { set_error("workspace too small"); return MMIF_E_WORKSPACE; }
    FwdLaunch L;
    L.win = WIN; L.sigma = 1.5; L.epi = EPI_SSIM; L.finalize = FIN_LOSS; L.do_sobel = 1;
    L.data_range = cfg->data_range; L.cfg = *cfg;
    double* sums = (double*)((unsigned char*)ws + fws);
    return launch_moment_fwd(L, i1, i2, f, B, H, W, sums, 8, out, ws, fws, (cudaStream_t)stream);
}

extern "C" int mmif_fusion_loss_bwd(const float* i1, const float* i2, const float* f, int B, int H, int W,
                                    const MmifLossCfg* cfg, const float* gout3, float* dF, void* ws, size_t ws_bytes,
                                    void* stream) {
    (void)ws; (void)ws_bytes;
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    rc = check_cfg(cfg);
    if (rc) return rc;
    if (!gout3 || !dF) { set_error("null gout3/dF"); return MMIF_E_NULL; }
    if (((uintptr_t)dF) & 3) { set_error("dF must be 4-byte aligned"); return MMIF_E_ALIGN; }
    const BwdGeom g = bwd_geom(B, H, W);
    BwdParams p;
    memset(&p, 0, sizeof(p));
    p.x1 = i1; p.x2 = i2; p.y = f; p.dF = dF; p.gout = gout3;
    p.B = B; p.H = H; p.W = W; p.Hout = g.Hout; p.Wout = g.Wout;
    p.seg_rows = g.seg_rows; p.nseg = g.nseg; p.nstrip = g.nstrip;
    make_taps(&p.taps, WIN, 1.5);
    const double L = cfg->data_range;
    p.C1 = (float)((0.01 * L) * (0.01 * L)); p.C2 = (float)((0.03 * L) * (0.03 * L));
    p.pixel_combine = cfg->pixel_combine; p.grad_combine = cfg->grad_combine;
    p.pixel_norm = cfg->pixel_norm; p.grad_norm = cfg->grad_norm;
    p.w_ssim = cfg->w_ssim; p.w_pixel = cfg->w_pixel; p.w_grad = cfg->w_grad;
    p.vec_store = ((W & 3) == 0) && ((((uintptr_t)dF) & 15) == 0);
    CUtensorMap m1, m2, my;
    p.use_tma = make_tensor_map(&m1, i1, B, H, W, kTWI, kRB) && make_tensor_map(&m2, i2, B, H, W, kTWI, kRB) &&
                make_tensor_map(&my, f, B, H, W, kTWI, kRB);
    if (!p.use_tma) { memset(&m1, 0, sizeof(m1)); memset(&m2, 0, sizeof(m2)); memset(&my, 0, sizeof(my)); }
    static bool attr_done = false;
    if (!attr_done) {
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemBwd)));
        attr_done = true;
    }
    dim3 grid(g.nstrip, g.nseg, B);
    fusion_loss_bwd_kernel<<<grid, kNT, sizeof(SmemBwd), (cudaStream_t)stream>>>(m1, m2, my, p);
    MMIF_CUDA(cudaGetLastError());
    return MMIF_OK;
}
