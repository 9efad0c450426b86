// Strip-streaming stencil engine shared by the loss kernels and the SSIM / VIF metric kernels.
//
// One CTA (128 threads) owns a strip of kTWI = 128 input columns and marches down a segment of
// rows in batches of kRB = 8 output rows:
//   * input rows arrive in 8-row groups through a 4-slot shared-memory ring (32 rows x 128
//     columns x 3 images = 48 KB), filled by TMA (one elected thread, cp.async.bulk.tensor +
//     mbarrier complete_tx) one group ahead of the compute; a plain-load fallback covers tensors
//     TMA cannot describe (row pitch not a multiple of 16 B);
//   * V-pass: thread = input column; the WIN-tap vertical blur of the packed map pairs
//     (x1,x2) (x1^2,x2^2) (x1*y,x2*y) (y,y^2) for 8 output rows is accumulated in registers with
//     FFMA2 and written to `vbuf`;
//   * H-pass: thread = (row, 8-column group); the horizontal WIN-tap blur reads `vbuf` with
//     conflict-free LDS.128 (lanes 0-7 of a quarter warp are 8 rows, row pitch = 16 mod 128 B)
//     and hands 8 pixels x 4 map pairs to a kernel-specific epilogue, still in registers.
// Moments are taken of mean-shifted data (x - c, c = one pixel of the CTA's tile) so that
// E[x^2] - mu^2 does not cancel catastrophically in fp32; the reference's value is recovered
// exactly (its 2-D fp32 window sums to 1 + eps, not 1) by the correction terms in `Shift`.
#pragma once
#include "common.cuh"

namespace mmif {

constexpr int kNT = 128;        // threads per CTA
constexpr int kTWI = 128;       // input columns per strip
constexpr int kRB = 8;          // output rows per batch
constexpr int kVCols = 136;     // vbuf columns (128 + read-ahead pad)
constexpr int kVPitch = 4 * kVCols + 2;   // float2 units per output row (pad 16 B -> pitch = 16 mod 128 B)

// SLOTS 8-row groups of RP columns per image.  Forward: 3 slots x 128 (36 KB; the group for batch b+1
// is fetched while batch b is in its H-pass) so that 3 CTAs fit an SM; backward: 4 slots x 132
// (row pitch = 16 mod 128 B: its combine step reads 8 rows at a time with LDS.128, conflict-free).
template <int SLOTS, int RP>
struct SmemT {
    static constexpr int kSlots = SLOTS;
    static constexpr int kRows = SLOTS * kRB;
    static constexpr int kPitch = RP;
    static constexpr int kGroupBytes = 3 * kRB * RP * 4;
    float ring[3][SLOTS * kRB][RP];
    alignas(16) float2 vbuf[kRB * kVPitch];           // 34944 B
    unsigned long long mbar[4];
    double red[8 * (kNT / 32)];
    int flag;
    __device__ __forceinline__ static int wrap(int r) {          // r in [0, 2*kRows)
        if (SLOTS == 4) return r & (kRows - 1);
        return r >= kRows ? r - kRows : r;
    }
    __device__ __forceinline__ static int wrap_any(int r) {      // any r >= 0
        if (SLOTS == 4) return r & (kRows - 1);
        return r % kRows;
    }
};

// ---- input ring -------------------------------------------------------------------------
struct RingSrc {
    const float* img[3];   // sample base pointers (generic path)
    int H, W;
    int row0, col0;        // image coordinates of ring-local (row 0, col 0)
    int n;                 // sample index (TMA coordinate 2)
    bool use_tma;
};

template <class SM>
__device__ __forceinline__ void ring_issue(SM& sm, const RingSrc& s, const CUtensorMap* m0, const CUtensorMap* m1,
                                           const CUtensorMap* m2, int g) {
    const int slot = g % SM::kSlots;
    if (s.use_tma) {
        if (threadIdx.x == 0) {
            unsigned long long* bar = &sm.mbar[slot];
            mbar_expect_tx((uint64_t*)bar, SM::kGroupBytes);
            tma_load_3d(&sm.ring[0][slot * kRB][0], m0, s.col0, s.row0 + g * kRB, s.n, (uint64_t*)bar);
            tma_load_3d(&sm.ring[1][slot * kRB][0], m1, s.col0, s.row0 + g * kRB, s.n, (uint64_t*)bar);
            tma_load_3d(&sm.ring[2][slot * kRB][0], m2, s.col0, s.row0 + g * kRB, s.n, (uint64_t*)bar);
        }
    } else {
        for (int tc = threadIdx.x; tc < SM::kPitch; tc += kNT) {
            const int col = s.col0 + tc;
            const bool cok = (col >= 0) && (col < s.W);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
#pragma unroll
                for (int r = 0; r < kRB; ++r) {
                    const int row = s.row0 + g * kRB + r;
                    float v = 0.f;
                    if (cok && row >= 0 && row < s.H) v = __ldg(s.img[k] + (size_t)row * s.W + col);
                    sm.ring[k][slot * kRB + r][tc] = v;
                }
            }
        }
    }
}
template <class SM>
__device__ __forceinline__ void ring_wait(SM& sm, const RingSrc& s, int g) {
    if (s.use_tma) mbar_wait((uint64_t*)&sm.mbar[g % SM::kSlots], (g / SM::kSlots) & 1);
}

// ---- mean shift + window-sum correction ---------------------------------------------------
// With x' = x - c and the reference window summing to s = 1 + eps:
//   mu_ref    = mu' + s c
//   var_ref   = E'[x'^2] - mu'^2 - eps c (2 mu' + c s)
//   cov_ref   = E'[x'y'] - mu_x' mu_y' - eps (c_x mu_y' + c_y mu_x' + c_x c_y s)
struct Shift {
    float2 c;        // (c1, c2) shifts of the two sources
    float cy;        // shift of the fused image
    float2 sc;       // s * c
    float scy;
    float2 k1;       // 2 eps c
    float2 k0;       // eps c^2 s
    float k1y, k0y;
    float2 ec;       // eps c            (multiplies mu_y')
    float2 nec;      // -eps c
    float ecy;       // eps cy           (multiplies mu_k')
    float2 eccs;     // eps c cy s
    float rho;       // window-sum mismatch of the separable taps (Taps::wrho)
};
__device__ __forceinline__ Shift make_shift(float c1, float c2, float cy, float s, float eps, float rho) {
    Shift h;
    h.rho = rho;
    h.c = f2(c1, c2);
    h.cy = cy;
    h.sc = f2(s * c1, s * c2);
    h.scy = s * cy;
    h.k1 = f2(2.f * eps * c1, 2.f * eps * c2);
    h.k0 = f2(eps * c1 * c1 * s, eps * c2 * c2 * s);
    h.k1y = 2.f * eps * cy;
    h.k0y = eps * cy * cy * s;
    h.ec = f2(eps * c1, eps * c2);
    h.nec = f2(-eps * c1, -eps * c2);
    h.ecy = eps * cy;
    h.eccs = f2(eps * c1 * cy * s, eps * c2 * cy * s);
    return h;
}

// Blurred shifted moments of one window position, as the H-pass delivers them.
struct Moments {
    float2 mk;    // (mu1', mu2')
    float2 ekk;   // (E[x1'^2], E[x2'^2])
    float2 eky;   // (E[x1'y'], E[x2'y'])
    float my;     // mu_y'
    float eyy;    // E[y'^2]
};
struct Stats {    // reference-equivalent statistics
    float2 mu;    // unshifted means of the sources
    float muy;
    float2 vk;    // variances before the clamp
    float vy;
    float2 cov;
};
__device__ __forceinline__ Stats stats_from(const Moments& m, const Shift& h) {
    Stats s;
    const float2 nrho = bcast(-h.rho), neg1 = bcast(-1.f);
    s.mu = add2(m.mk, h.sc);
    s.muy = m.my + h.scy;
    // vk = ekk - t - rho t, t = mk (mk + k1) + k0   [rho (ekk - 2 mk^2) = rho (ekk - t) - rho t; the first part is < 0.1 ulp
    // of ekk - t and dropped, the second is what decides the sign of a flat region's variance]
    const float2 t = fma2(m.mk, add2(m.mk, h.k1), h.k0);
    s.vk = fma2(nrho, t, fma2(t, neg1, m.ekk));
    const float ty = fmaf(m.my, m.my + h.k1y, h.k0y);
    s.vy = fmaf(-h.rho, ty, m.eyy - ty);
    // cov = eky - u - rho u, u = mk (my + ecy) + ec my + eccs
    const float2 u = fma2(m.mk, bcast(m.my + h.ecy), fma2(h.ec, bcast(m.my), h.eccs));
    s.cov = fma2(nrho, u, fma2(u, neg1, m.eky));
    return s;
}

// ---- V-pass: vertical WIN-tap blur of the four packed moment maps --------------------------
// `base` = ring-local index of the first input row of this batch (multiple of 8); thread t blurs ring
// column t + coff and writes vbuf column t.
template <int WIN, class SM>
__device__ __forceinline__ void vpass_moments(SM& sm, const Taps& tp, const Shift& h, int base, int coff = 0) {
    const int t = threadIdx.x, tr = t + coff;
    float2 acc[kRB][4];
    const float2 negc = f2(-h.c.x, -h.c.y);
    // the kRB + WIN - 1 input rows lie in consecutive 8-row ring slots: one base pointer per slot, constant row offsets inside
    // it (no wrap arithmetic per row: every instruction costs a dispatch-port cycle)
    constexpr int kSeg = (kRB + WIN - 1 + kRB - 1) / kRB;
    constexpr int kPlane = SM::kRows * SM::kPitch;
    const float* seg[kSeg];
#pragma unroll
    for (int k = 0; k < kSeg; ++k) seg[k] = &sm.ring[0][SM::wrap(base + k * kRB)][tr];
#pragma unroll
    for (int rr = 0; rr < kRB + WIN - 1; ++rr) {
        const float* rp = seg[rr / kRB] + (rr % kRB) * SM::kPitch;
        const float y = rp[2 * kPlane] - h.cy;
        float2 P[4];
        P[0] = add2(f2(rp[0], rp[kPlane]), negc);
        P[1] = mul2(P[0], P[0]);
        P[2] = muls(y, P[0]);
        P[3] = f2(y, y * y);
#pragma unroll
        for (int o = 0; o < kRB; ++o) {
            const int k = rr - o;
            if (k >= 0 && k < WIN) {
#pragma unroll
                for (int m = 0; m < 4; ++m) acc[o][m] = (k == 0) ? muls(tp.w[0], P[m]) : fmas(tp.w[k], P[m], acc[o][m]);
            }
        }
        if (rr >= WIN - 1) {
            const int o = rr - (WIN - 1);
#pragma unroll
            for (int m = 0; m < 4; ++m) sm.vbuf[o * kVPitch + m * kVCols + t] = acc[o][m];
        }
    }
}

// ---- H-pass: horizontal WIN-tap blur, 8 pixels x NM packed maps per thread -------------------
// src row pointer = buffer + o*pitch + 8*g (float2 units); map stride `mstride`.
// REV: use taps reversed (adjoint pass).  Results in acc[j][m].
template <int WIN, int NM, bool REV, int NOUT = 8>
__device__ __forceinline__ void hpass(const float2* __restrict__ src, int mstride, const Taps& tp, float2 (&acc)[NOUT][NM]) {
#pragma unroll
    for (int kk = 0; kk < NOUT + WIN - 1; kk += 2) {
        float4 q[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) q[m] = *reinterpret_cast<const float4*>(src + m * mstride + kk);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int col = kk + half;
#pragma unroll
            for (int j = 0; j < NOUT; ++j) {
                const int k = col - j;
                if (k >= 0 && k < WIN) {
                    const float w = REV ? tp.w[WIN - 1 - k] : tp.w[k];
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        const float2 v = half ? f2(q[m].z, q[m].w) : f2(q[m].x, q[m].y);
                        acc[j][m] = (k == 0) ? muls(w, v) : fmas(w, v, acc[j][m]);
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ Moments moments_of(const float2 (&a)[4]) {
    Moments m;
    m.mk = a[0];
    m.ekk = a[1];
    m.eky = a[2];
    m.my = a[3].x;
    m.eyy = a[3].y;
    return m;
}

// Tile shift constants, per image: the minimum of a 16 x 8 grid of samples of the CTA's region
// [r0, r0+nr) x [c0, c0+128).  For non-negative data 0 <= c <= (most) v, so |v - c| <= |v|: the
// rounding noise of the shifted moments (~1e-7 (v-c)^2) never exceeds that of the reference's
// unshifted fp32 moments (~1e-7 v^2), dark flat regions (v == c) become exact, and the values of a
// well-exposed tile shrink by its floor.  Must be called by all threads (one __syncthreads).
template <class SM>
__device__ __forceinline__ Shift tile_shift(SM& sm, const float* x1, const float* x2, const float* y, int H, int W, int r0, int nr,
                                            int c0, const Taps& tp) {
    const int t = threadIdx.x;
    int r = r0 + ((t >> 4) * nr) / 8 + nr / 16;
    int c = c0 + (t & 15) * 8 + 4;
    r = min(max(r, 0), H - 1);
    c = min(max(c, 0), W - 1);
    const size_t off = (size_t)r * W + c;
    float v[3] = {__ldg(x1 + off), __ldg(x2 + off), __ldg(y + off)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!(fabsf(v[k]) <= 3.0e38f)) v[k] = 3.0e38f;       // ignore NaN / inf samples
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    }
    float* scratch = reinterpret_cast<float*>(sm.red);
    if ((t & 31) == 0) { scratch[(t >> 5) * 3 + 0] = v[0]; scratch[(t >> 5) * 3 + 1] = v[1]; scratch[(t >> 5) * 3 + 2] = v[2]; }
    __syncthreads();
    float c3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float m = scratch[k];
#pragma unroll
        for (int w = 1; w < kNT / 32; ++w) m = fminf(m, scratch[w * 3 + k]);
        c3[k] = (m < 3.0e38f) ? m : 0.f;
    }
    __syncthreads();
    return make_shift(c3[0], c3[1], c3[2], tp.wsum, tp.weps, tp.wrho);
}

__device__ __forceinline__ float norm_val(float d, int norm) { return norm == MMIF_NORM_L1 ? fabsf(d) : d * d; }
__device__ __forceinline__ float sgn(float d) { return (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); }
__device__ __forceinline__ float norm_der(float d, int norm) { return norm == MMIF_NORM_L1 ? sgn(d) : 2.f * d; }

}  // namespace mmif
