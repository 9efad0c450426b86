// Fusion objective (reference core/loss.py): SSIM('ssim') + PixelLoss + GradLoss, forward and
// backward, as two strip-streaming kernels (see stencil.cuh for the engine).
//
//   fusion_loss_fwd : reads I1, I2, If once (12 B/pixel), writes per-CTA partial sums only; the last
//                     CTA of each sample reduces that sample in a fixed order (deterministic), the
//                     last sample-finisher writes the loss scalars.  One launch.
//   fusion_loss_bwd : recomputes the moments, forms the per-window derivative coefficients, runs the
//                     adjoint blur and the folded Sobel adjoint and writes dL/dIf once
//                     (12 B read + 4 B written per pixel).  One launch.
#include <stdlib.h>
#include <mutex>
#include <queue>
#include <vector>
#include "moment_fwd.cuh"

namespace mmif {

constexpr int WIN11 = 11;
// Backward tile geometry for a WIN-tap window: the ring origin is j0 - OFF (a multiple of 4 columns,
// >= HALO, so that the float4 reads of the combine step stay aligned); the 128 V-pass threads take ring
// columns VOFF .. VOFF+127 (the ring rows are 132 wide), i.e. image columns from j0 - HALO, so the WC =
// 128 - HALO window columns start exactly HALO left of the gradient strip and TG = 128 - 2 HALO gradient
// columns (rounded down to a multiple of 4) come out of every strip.  WIN = 11: OFF 12, VOFF 2, TG 108, WC 118.
template <int WIN>
struct BG {
    static constexpr int HALO = WIN - 1;
    static constexpr int OFF = (HALO + 3) / 4 * 4;
    static constexpr int VOFF = OFF - HALO;
    static constexpr int TG = (kTWI - 2 * HALO) / 4 * 4;
    static constexpr int WC = kTWI - HALO;
};
static int bwd_tg(int win) { const int halo = win - 1; return (kTWI - 2 * halo) / 4 * 4; }

// Row segments of the backward / single-pass grid: the first n_tall segments are seg_rows high, the rest seg_short
// (both multiples of 8; seg_short == seg_rows = the uniform scheme).  The grid is (strip, sample, segment) with the segment
// index SLOWEST, so the hardware dispatcher hands out all tall segments of all samples first and the short ones last:
// longest-processing-time-first list scheduling, which trims the partial last wave that costs a per-rank batch of 8
// (304 columns on 296 CTA slots) 8 % in the uniform scheme.  Every sample and strip is cut the same way, so a sample's
// result does not depend on its position in the batch.
// Warp-specialised kernel only: the LAST fine_strips strips of every sample form a second class that is cut into short
// segments of fine_rows rows and dispatched after everything else.  With one CTA per SM the other strips can then be
// full-height columns in whole waves (W = 4096: 37 of the 38 strips x B columns = B / 4 waves of 148) and the fine class
// fills the SMs up to a common finish line; every sample is still cut the same way.
struct BwdGeom { int Hout, Wout, seg_rows, seg_short, n_tall, nseg, nstrip, fine_strips, fine_rows, nseg_fine; };
static int geom_ctas_per_sample(const BwdGeom& g) { return (g.nstrip - g.fine_strips) * g.nseg + g.fine_strips * g.nseg_fine; }

static int geom_nseg(int H, int T, int n_tall, int s) {
    const int tall_rows = n_tall * T;
    return tall_rows >= H ? ceil_div(H, T) : n_tall + ceil_div(H - tall_rows, s);
}

// Makespan (in row units) of `cols` columns cut into the given segments on 148 SMs x 2 slots, CTAs dispatched in launch
// order to the first slot that frees; a CTA alone on its SM runs kSolo times faster (measured 1.39).  Event simulation.
static double simulate_makespan(int H, int cols, int T, int n_tall, int s, int extra) {
    constexpr int kSM = 148;
    constexpr double kSolo = 1.39;
    const int nseg = geom_nseg(H, T, n_tall, s);
    const long long total = (long long)cols * nseg;
    auto work_of = [&](long long idx) {          // launch order: segment slowest
        const int seg = (int)(idx / cols);
        int i0, i1;
        if (seg < n_tall) { i0 = seg * T; i1 = i0 + T; } else { i0 = n_tall * T + (seg - n_tall) * s; i1 = i0 + s; }
        if (i1 > H) i1 = H;
        return (double)(i1 - i0 + extra);
    };
    double rem[kSM][2];
    double now[kSM];
    unsigned ver[kSM];
    long long next = 0;
    for (int k = 0; k < kSM; ++k) { rem[k][0] = rem[k][1] = 0.0; now[k] = 0.0; ver[k] = 0u; }
    for (int slot = 0; slot < 2; ++slot)
        for (int k = 0; k < kSM && next < total; ++k) rem[k][slot] = work_of(next++);
    struct Ev { double t; int sm; int slot; unsigned ver; };
    auto later = [](const Ev& a, const Ev& b) { return a.t > b.t; };
    std::priority_queue<Ev, std::vector<Ev>, decltype(later)> pq(later);
    auto push_next = [&](int k) {               // earliest completion on SM k under its current pairing
        const bool a0 = rem[k][0] > 0.0, a1 = rem[k][1] > 0.0;
        if (!a0 && !a1) return;
        const double rate = (a0 && a1) ? 1.0 : kSolo;
        const int slot = (a0 && (!a1 || rem[k][0] <= rem[k][1])) ? 0 : 1;
        pq.push(Ev{now[k] + rem[k][slot] / rate, k, slot, ver[k]});
    };
    for (int k = 0; k < kSM; ++k) push_next(k);
    double makespan = 0.0;
    while (!pq.empty()) {
        const Ev e = pq.top();
        pq.pop();
        if (e.ver != ver[e.sm]) continue;
        const int k = e.sm;
        const bool both = rem[k][0] > 0.0 && rem[k][1] > 0.0;
        const double rate = both ? 1.0 : kSolo;
        const double dt = e.t - now[k];
        for (int slot = 0; slot < 2; ++slot)
            if (rem[k][slot] > 0.0) rem[k][slot] -= dt * rate;
        rem[k][e.slot] = 0.0;
        now[k] = e.t;
        makespan = e.t;
        if (next < total) rem[k][e.slot] = work_of(next++);
        ++ver[k];
        push_next(k);
    }
    return makespan;
}

// The same for the warp-specialised kernel: ONE CTA per SM, so plain list scheduling on 148 machines.  The CTAs come in at most
// three classes of equal length dispatched in order (tall segments, short segments, the ragged last segment), so the machines
// are tracked as a handful of (load, count) groups: exact, and ~total / 148 steps instead of total heap operations.
struct ListSched {             // greedy list scheduling of runs of equal jobs on 148 machines, tracked as (load, count) groups
    struct Grp { double load; int cnt; };
    Grp grp[24];
    int ng;
    ListSched() : ng(1) { grp[0] = Grp{0.0, 148}; }
    void add(int n, double len) {
        while (n > 0) {
            int lo = 0;
            for (int k = 1; k < ng; ++k)
                if (grp[k].load < grp[lo].load) lo = k;
            const int take = grp[lo].cnt < n ? grp[lo].cnt : n;
            const double nl = grp[lo].load + len;
            grp[lo].cnt -= take;
            if (grp[lo].cnt == 0) grp[lo] = grp[--ng];
            int hit = -1;
            for (int k = 0; k < ng; ++k)
                if (grp[k].load == nl) hit = k;
            if (hit >= 0) grp[hit].cnt += take;
            else if (ng < 24) grp[ng++] = Grp{nl, take};
            else { grp[0].cnt += take; if (grp[0].load < nl) grp[0].load = nl; }      // cannot happen with a handful of classes
            n -= take;
        }
    }
    void add_segments(int H, int cols, int T, int n_tall, int s, int extra) {
        const int nseg = geom_nseg(H, T, n_tall, s);
        for (int seg = 0; seg < nseg; ++seg) {
            int i0, i1;
            if (seg < n_tall) { i0 = seg * T; i1 = i0 + T; } else { i0 = n_tall * T + (seg - n_tall) * s; i1 = i0 + s; }
            if (i1 > H) i1 = H;
            add(cols, (double)(i1 - i0 + extra));
        }
    }
    double makespan() const {
        double m = 0.0;
        for (int k = 0; k < ng; ++k)
            if (grp[k].cnt > 0 && grp[k].load > m) m = grp[k].load;
        return m;
    }
};
static double simulate_makespan_1(int H, int cols, int T, int n_tall, int s, int extra) {
    ListSched ls;
    ls.add_segments(H, cols, T, n_tall, s, extra);
    return ls.makespan();
}

// ws: geometry for fusion_loss_ws_kernel (one CTA per SM) instead of fusion_loss_bwd_kernel (two)
static BwdGeom bwd_geom(int B, int H, int W, int win = WIN11, bool ws = false) {
    BwdGeom g;
    g.Hout = H - (win - 1); g.Wout = W - (win - 1);
    g.nstrip = ceil_div(W, bwd_tg(win));
    // halo + batch rounding + prologue, in rows; the warp-specialised pipeline also fills and drains (3 stages)
    static const int ws_extra = getenv("MMIF_WS_EXTRA") ? atoi(getenv("MMIF_WS_EXTRA")) : 12;
    const int extra = 2 * (win - 1) + 8 + (ws ? ws_extra : 0);
    g.seg_rows = ws ? pick_seg_rows(H, B * g.nstrip, 148, extra, 0.0) : pick_seg_rows(H, B * g.nstrip, 2 * 148, extra, 0.72);
    g.seg_short = g.seg_rows; g.n_tall = ceil_div(H, g.seg_rows);
    g.nseg = g.n_tall;
    g.fine_strips = 0; g.fine_rows = g.seg_rows; g.nseg_fine = 0;
    // memo: the search below simulates a few dozen schedules (~1 ms); a training loop asks for one shape
    struct Key { int B, H, W, win; bool ws; BwdGeom g; };
    // process-wide (the backward runs on autograd's thread and must not repeat the search) and large enough for a training
    // loop that alternates a few shapes: one shape occupies six entries (five windows + the warp-specialised geometry)
    constexpr int kMemo = 96;
    static Key memo[kMemo];
    static int memo_n = 0, memo_next = 0;
    static std::mutex memo_mu;
    std::unique_lock<std::mutex> memo_lock(memo_mu);
    const bool forced = ws && getenv("MMIF_WS_GEOM") != nullptr;
    if (forced) {                // measurement aid: MMIF_WS_GEOM="T,n_tall,s" forces the segments (tools/ws_geom_force.py)
        int fT = 0, fn = 0, fs = 0, fF = 0, fr = 0;
        const int got = sscanf(getenv("MMIF_WS_GEOM"), "%d,%d,%d,%d,%d", &fT, &fn, &fs, &fF, &fr);
        if (got >= 3 && fT >= 16 && fs >= 16 && fT % 8 == 0 && fs % 8 == 0) {
            g.seg_rows = fT; g.seg_short = fs; g.nseg = geom_nseg(H, fT, fn, fs); g.n_tall = fn < g.nseg ? fn : g.nseg;
            if (got == 5 && fF >= 1 && fF < g.nstrip && fr >= 16 && fr % 8 == 0) { g.fine_strips = fF; g.fine_rows = fr; g.nseg_fine = ceil_div(H, fr); }
            return g;
        }
    }
    for (int i = 0; i < memo_n && !forced; ++i)
        if (memo[i].B == B && memo[i].H == H && memo[i].W == W && memo[i].win == win && memo[i].ws == ws) return memo[i].g;
    memo_lock.unlock();          // the search below takes milliseconds: run it unlocked (a duplicate entry is harmless)
    const int cols = B * g.nstrip;
    static const bool uniform_only = getenv("MMIF_UNIFORM_SEGMENTS") != nullptr;      // A/B switch for the measurements
    auto simulate = [&](int T, int n_tall, int s) {
        return ws ? simulate_makespan_1(H, cols, T, n_tall, s, extra) : simulate_makespan(H, cols, T, n_tall, s, extra);
    };
    if (ws && !uniform_only) {
        // one CTA per SM: the closed-form tail term of pick_seg_rows ranks the candidates wrongly (B = 64: one 3072-row segment,
        // 17 waves, instead of two of 1536, 33 waves; measured 15.03 vs 14.86 ms), while list scheduling simulated on 148 SMs
        // with 40 rows of per-CTA overhead reproduces every measured ordering (tools/ws_geom_scan.sh).  Full search: the number
        // of fine strips F (0..3), for the other strips tall heights H / k (k = 1..24), short heights T / {1, 2, 3, 4, 6, 8} and
        // every count of tall segments, for the fine strips every height H / k down to 32 rows.
        double best = 1e300;
        const int divs[6] = {1, 2, 3, 4, 6, 8};
        static const int max_fine = getenv("MMIF_WS_MAX_FINE") ? atoi(getenv("MMIF_WS_MAX_FINE")) : 3;
        for (int F = 0; F <= max_fine && F < g.nstrip; ++F) {
            const int colsA = B * (g.nstrip - F), colsF = B * F;
            for (int k = 1; k <= 24; ++k) {
                const int T = ceil_div(ceil_div(H, k), 8) * 8;
                if (T < 32 && k > 1) break;
                for (int di = 0; di < (F == 0 ? 6 : 1); ++di) {     // with a fine class the coarse strips stay uniform: it balances
                    const int sh = ceil_div(ceil_div(T, divs[di]), 8) * 8;
                    if (sh < 16 || (di > 0 && sh >= T)) continue;
                    const int full = ceil_div(H, T);
                    for (int nt = (di == 0 ? full : 0); nt <= full && (nt == full || nt * T < H); ++nt) {
                        const int nseg = geom_nseg(H, T, nt, sh);
                        if ((long long)colsA * nseg > 40000) continue;
                        ListSched base;
                        base.add_segments(H, colsA, T, nt, sh, extra);
                        if (F == 0) {
                            const double m = base.makespan();
                            if (m < best * (di == 0 ? 1.0 : 0.995)) {
                                best = m; g.seg_rows = T; g.seg_short = sh; g.n_tall = nt < nseg ? nt : nseg; g.nseg = nseg;
                                g.fine_strips = 0; g.fine_rows = T; g.nseg_fine = 0;
                            }
                            continue;
                        }
                        for (int kf = 1; kf <= 96; ++kf) {
                            const int fr = ceil_div(ceil_div(H, kf), 8) * 8;
                            if (fr < 32) break;
                            const int nsf = ceil_div(H, fr);
                            if (nsf != kf || (long long)colsF * nsf > 20000) continue;
                            ListSched ls = base;
                            ls.add_segments(H, colsF, fr, nsf, fr, extra);
                            const double m = ls.makespan();
                            if (m < best * 0.99) {             // a fine class has to earn its extra halo rows
                                best = m; g.seg_rows = T; g.seg_short = sh; g.n_tall = nt < nseg ? nt : nseg; g.nseg = nseg;
                                g.fine_strips = F; g.fine_rows = fr; g.nseg_fine = nsf;
                            }
                        }
                    }
                }
            }
        }
    } else if (!uniform_only && win == WIN11 && (long long)cols * g.nseg <= 40000 && g.seg_rows >= 32) {
        // (the extended SSIM-only launches with the 9/7/5/3-tap windows keep the closed-form uniform segments: their geometry is
        // also computed for every workspace sizing, and the event simulation below costs tens of milliseconds per window)
        double best = simulate(g.seg_rows, g.n_tall, g.seg_rows);
        const BwdGeom uni = g;
        const int talls[4] = {uni.seg_rows, (uni.seg_rows * 5 / 4 + 7) / 8 * 8, (uni.seg_rows * 3 / 2 + 7) / 8 * 8, uni.seg_rows * 2};
        const int divs[5] = {2, 3, 4, 6, 8};
        for (int ti = 0; ti < 4; ++ti) {
            const int T = talls[ti] > H ? (H + 7) / 8 * 8 : talls[ti];
            for (int di = 0; di < 5; ++di) {
                const int sh = (T / divs[di] + 7) / 8 * 8;
                if (sh < 16 || sh >= T) continue;
                const int full = H / T;
                for (int nt = (full > 3 ? full - 3 : 0); nt * T < H; ++nt) {
                    const int nseg = geom_nseg(H, T, nt, sh);
                    if ((long long)cols * nseg > 40000) continue;
                    const double m = simulate(T, nt, sh);
                    if (m < best * 0.985) { best = m; g.seg_rows = T; g.seg_short = sh; g.n_tall = nt; g.nseg = nseg; }
                }
            }
        }
    }
    static const bool debug_geom = getenv("MMIF_DEBUG_GEOM") != nullptr;
    if (debug_geom)
        fprintf(stderr, "[mmif] bwd geometry%s B=%d H=%d W=%d win=%d: %d strips, %d segments = %d x %d rows + %d x %d rows\n", ws ? " (ws)" : "", B, H, W, win,
                g.nstrip, g.nseg, g.n_tall < g.nseg ? g.n_tall : g.nseg, g.seg_rows, g.nseg > g.n_tall ? g.nseg - g.n_tall : 0, g.seg_short);
    if (debug_geom && g.fine_strips)
        fprintf(stderr, "[mmif]   fine class: the last %d strip(s) of every sample in %d segments of %d rows\n", g.fine_strips, g.nseg_fine, g.fine_rows);
    memo_lock.lock();
    Key& k = memo[memo_next];
    k.B = B; k.H = H; k.W = W; k.win = win; k.ws = ws; k.g = g;
    memo_next = (memo_next + 1) % kMemo;
    if (memo_n < kMemo) ++memo_n;
    return g;
}

// =============================================================================== backward
struct BwdParams {
    const float* x1; const float* x2; const float* y;
    float* dF;
    const float* gout[3];    // upstream gradients of the three loss values (device scalars; all NULL = unit upstream;
                             // gout[0] set and gout[1] / gout[2] NULL = those two terms get no gradient)
    int B, H, W, Hout, Wout;
    int seg_rows, seg_short, n_tall, nseg, nstrip;   // row segments: see BwdGeom
    int fine_strips, fine_rows, nseg_fine;           // warp-specialised kernel: the fine class (BwdGeom)
    Taps taps;
    float C1, C2;
    int pixel_combine, grad_combine, pixel_norm, grad_norm;
    float w_ssim, w_pixel, w_grad;
    int use_tma;
    int vec_store;
    const float* pair_w;     // nullptr or [B][2]: per-sample weights of the two SSIM pairs ('w-ssim', MS-SSIM levels)
    float ssim_base;         // k_ssim = g[0] * ssim_base / (Hout*Wout); default w_ssim * (-0.5) / B
    int cs_only;             // differentiate the contrast-structure term only (MS-SSIM levels 0..3, loss.py:144-145)
    int msw;                 // MSW_SSIM (loss.py:226-237): per-window pair weights gamma, 1 - gamma from the source variances
    int accum;               // dF += instead of dF =
    int do_sobel;            // 0: SSIM term only (no pixel / Sobel adjoint)
    const float* dF_unit;    // != nullptr: gradient already computed for unit upstream (single-pass forward);
                             // if the three upstream gradients are equal the kernel only rescales it
    FinishParams fin;        // ZMODE: loss sums / finish (same protocol as the forward kernel)
};

constexpr int kCPitch = 2 * kTWI + 2;   // float2 units per coefficient row (2 pair-maps x 128 + 16 B pad)
constexpr int kTMC = 128;               // gbuf columns (>= the widest gradient strip: 124 for WIN = 3)

constexpr int kRPB = kTWI + 4;           // backward ring row pitch (132 floats = 16 mod 128 B)
using SmemB = SmemT<4, kRPB>;

// r * sign(g) with sign(0) = 0 (torch.sign / autograd of abs at 0)
__device__ __forceinline__ float mulsign(float r, float g) {
    return (g == 0.f) ? 0.f : __int_as_float(__float_as_int(r) ^ (__float_as_int(g) & 0x80000000));
}

struct SmemBwd {
    SmemB s;                             // ring + vbuf (vbuf also hosts the B1->B2 buffer)
    alignas(16) float2 cbuf[kRB * kCPitch];          // coefficient rows of the current batch: (a,b) and (c1,c2)
    alignas(16) float gbuf[kRB][kTMC + 4];   // pixel + Sobel gradient of the batch rows (row pitch 528 B = 16 mod 128)
};

// FAST : mode='max' + 'l1' for both terms (train.py:67-68,307-308), no run-time mode switches.
// ZMODE: single-pass variant — the same launch also accumulates the three loss values (every window
//        position / pixel is owned by exactly one CTA: the one whose gradient tile contains it), so
//        forward + backward cost one kernel and 16 B/pixel instead of two kernels and 28 B/pixel.
// EXT  : the extended SSIM-only features (per-sample / per-position pair weights, cs-only, accumulate)
//        used by 'w-ssim', MS-SSIM and MSW-SSIM; compiled out of the training-path instantiations.
template <int WIN, bool FAST, bool ZMODE, bool EXT>
__global__ void __launch_bounds__(kNT, 2)
fusion_loss_bwd_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                       const __grid_constant__ CUtensorMap mapy, const BwdParams p) {
    constexpr int HALO = BG<WIN>::HALO, kOFF = BG<WIN>::OFF, kVOFF = BG<WIN>::VOFF, kTG = BG<WIN>::TG, kWC = BG<WIN>::WC;
    const bool unit_up = (p.gout[0] == nullptr) && (p.gout[1] == nullptr) && (p.gout[2] == nullptr);
    const float g_ssim = unit_up ? 1.f : (p.gout[0] ? __ldg(p.gout[0]) : 0.f);
    const float g_pix = unit_up ? 1.f : ((p.gout[1] && p.do_sobel) ? __ldg(p.gout[1]) : 0.f);
    const float g_grad = unit_up ? 1.f : ((p.gout[2] && p.do_sobel) ? __ldg(p.gout[2]) : 0.f);
    if (!ZMODE && p.dF_unit != nullptr && g_ssim == g_pix && g_pix == g_grad) {
        // total = l1 + l2 + l3 (train.py:69): the upstream gradients are one common scalar, and the
        // single-pass forward already produced d(total)/dIf for unit upstream: rescale_unit_kernel
        // (launched just before this kernel) has rescaled it, nothing to recompute.
        return;
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemBwd& sb = *reinterpret_cast<SmemBwd*>(smem_raw);
    SmemB& sm = sb.s;
    const int strip = blockIdx.x, n = blockIdx.y, seg = blockIdx.z;        // segment slowest: tall segments are dispatched first
    const int j0 = strip * kTG;
    const int i0 = (seg < p.n_tall) ? seg * p.seg_rows : p.n_tall * p.seg_rows + (seg - p.n_tall) * p.seg_short;
    const int seg_h = (seg < p.n_tall) ? p.seg_rows : p.seg_short;
    const int jw0 = j0 - kOFF;             // first input column of the ring (multiple of 4); windows start at jw0 + kVOFF
    const int R0 = i0 - HALO;              // first window / input row of the segment
    const int iend = min(i0 + seg_h, p.H);
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
    const float npx = (float)p.B * (float)p.H * (float)p.W;
    const float k_ssim2 = 2.f * g_ssim * p.ssim_base / ((float)p.Hout * (float)p.Wout);   // the blurred coefficients are halved
    const float2 pairw = (EXT && p.pair_w) ? f2(__ldg(p.pair_w + 2 * n), __ldg(p.pair_w + 2 * n + 1)) : f2(1.f, 1.f);
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
    float2* tbuf = sm.vbuf;                           // [8][2][128]+pad, alive between B1 and B2
    float2 carry[HALO][2];                            // vertical adjoint state: pending gradient rows
#pragma unroll
    for (int d = 0; d < HALO; ++d) carry[d][0] = carry[d][1] = f2(0.f, 0.f);

    if (EXT || !p.do_sobel) {
        for (int i = t; i < kRB * (kTMC + 4); i += kNT) (&sb.gbuf[0][0])[i] = 0.f;     // B2 adds gbuf unconditionally
    }
    // Sobel-adjoint phase: column of this thread and its sliding state
    const int s_ci = 30 * warp + lane - 1;            // column relative to j0: -1 .. 120 (lanes 0 / 31 are halo lanes)
    const int s_c = j0 + s_ci;
    const bool s_colok = (s_c >= 0) && (s_c < p.W) && (s_ci <= kTG);
    const bool s_own = (lane >= 1) && (lane <= 30) && (s_ci >= 0) && (s_ci < kTG) && (s_c < jend);
    const int s_cc = min(max(s_c, 0), p.W - 1);
    const int s_t0 = min(s_cc - jw0, kRPB - 1);
    const int s_tm = min(((s_cc == 0) ? 1 : s_cc - 1) - jw0, kRPB - 1);
    const int s_tp = min(((s_cc == p.W - 1) ? p.W - 2 : s_cc + 1) - jw0, kRPB - 1);
    // fast-path constants: ring pointer of the thread's column, ownership as a factor
    constexpr int kRingImg = SmemB::kRows * kRPB;
    // (strips whose Sobel columns j0-1 .. j0+kTG all lie in [2, W-3]: no column reflection or folding)
    const bool s_strip_int = (j0 >= 3) && (j0 + kTG <= p.W - 3);
    const float* s_p0 = &sm.ring[0][0][min(max(s_c - jw0, 1), kRPB - 2)];
    const float s_ownf = s_own ? 1.f : 0.f;
    float* s_gdst = &sb.gbuf[0][s_own ? s_ci : 0];
    float2 s_dA = f2(0.f, 0.f), s_dB = s_dA, s_sA = s_dA, s_sB = s_dA, s_ucA = s_dA, s_ucB = s_dA;
    float s_dAy = 0.f, s_dBy = 0.f, s_sAy = 0.f, s_sBy = 0.f, s_ucAy = 0.f, s_ucBy = 0.f;
    float s_hxA = 0.f, s_hxB = 0.f, s_vyA = 0.f, s_vyB = 0.f;
    float2 z_ss = f2(0.f, 0.f), z_cs = z_ss, z_sg = z_ss;     // ZMODE loss accumulators
    float z_pix = 0.f, z_grad = 0.f;

    // window columns of this thread's H-pass group: bit j of a_vmask = column carries a valid window,
    // of a_zmask = ... and its SSIM value belongs to this CTA's tile (single-pass loss sums)
    unsigned a_vmask = 0u, a_zmask = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int pw = hg * 8 + j, pc = jw0 + kVOFF + pw;
        const bool v = (pw < kWC) && (pc >= 0) && (pc < p.Wout);
        a_vmask |= v ? (1u << j) : 0u;
        a_zmask |= (v && pc >= j0 && pc < jend) ? (1u << j) : 0u;
    }
    for (int b = 0; b < nb; ++b) {
        const int Rb = R0 + b * kRB;                  // first gradient row of this batch
        ring_wait(sm, src, b + 2);
        const bool emit = (Rb + kRB > i0);            // any owned gradient row in this batch
        if (b + 1 < nb) ring_issue(sm, src, &map1, &map2, &mapy, b + 3);   // slot of group b-1: nobody reads it any more
        // ---------------- S: Sobel + pixel adjoint of gradient rows [Rb, Rb+8) -> gbuf ------------
        // thread = column (warps own 30 columns + 1 halo lane each side); one input row per step:
        // input row q' -> Sobel / tx,ty of row q'-1 -> (neighbour columns by shuffle) -> G of row q'-2.
        // All sliding state lives in registers across batches, so every input row is visited once.
        // Interior batches (no row reflection / folding, every row owned, 'max' + 'l1'): branch-free,
        // fully unrolled — the sliding state is renamed instead of moved, borders cost nothing.
        const bool s_fast = FAST && !EXT && p.do_sobel && s_strip_int && (Rb >= max(2, i0)) && (Rb + 9 <= min(iend, p.H - 1));
        if (s_fast) {
            // Three passes over the 8 rows so that the 8 independent row chains overlap (the phase is
            // latency-bound, not issue-bound): 1. loads, Sobel, sign factors; 2. neighbour exchange;
            // 3. vertical combination and store.
            const int bb = (b & 3) * kRB + 2;
            float tx[kRB], ty[kRB], pg[kRB];
#pragma unroll
            for (int step = 0; step < kRB; ++step) {
                const int lro = ((bb + step) & (SmemB::kRows - 1)) * kRPB;      // ring row of input row Rb + 2 + step
                const float* rp = s_p0 + lro;
                const float2 um = f2(rp[-1], rp[kRingImg - 1]);
                const float2 uc = f2(rp[0], rp[kRingImg]);
                const float2 up = f2(rp[1], rp[kRingImg + 1]);
                const float umy = rp[2 * kRingImg - 1], ucy = rp[2 * kRingImg], upy = rp[2 * kRingImg + 1];
                const float2 d = fma2(bcast(-1.f), um, up);
                const float2 sv = fma2(bcast(2.f), uc, add2(um, up));
                const float2 gx = fma2(bcast(2.f), s_dB, add2(s_dA, d));   // Sobel of row Rb + 1 + step
                const float2 gy = fma2(bcast(-1.f), s_sA, sv);
                const float dy = upy - umy;
                const float sy = fmaf(2.f, ucy, umy + upy);
                const float gxy = fmaf(2.f, s_dBy, s_dAy + dy);
                const float gyy = sy - s_sAy;
                const float S1 = fabsf(gx.x) + fabsf(gy.x), S2 = fabsf(gx.y) + fabsf(gy.y), Sy = fabsf(gxy) + fabsf(gyy);
                const float D = Sy - fmaxf(S1, S2);
                const float r = mulsign(k_grad, D);
                tx[step] = mulsign(r, gxy);
                ty[step] = mulsign(r, gyy);
                const float dp = s_ucAy - fmaxf(s_ucA.x, s_ucA.y);         // pixel term of row Rb + step
                pg[step] = mulsign(k_pix, dp);
                if (ZMODE) {
                    z_grad = fmaf(s_ownf, fabsf(D), z_grad);
                    z_pix = fmaf(s_ownf, fabsf(dp), z_pix);
                }
                s_dA = s_dB; s_dB = d; s_sA = s_sB; s_sB = sv; s_dAy = s_dBy; s_dBy = dy; s_sAy = s_sBy; s_sBy = sy;
                s_ucA = s_ucB; s_ucB = uc; s_ucAy = s_ucBy; s_ucBy = ucy;
            }
#pragma unroll
            for (int step = 0; step < kRB; ++step) {
                const float txl = __shfl_up_sync(0xffffffffu, tx[step], 1), txr = __shfl_down_sync(0xffffffffu, tx[step], 1);
                const float tyl = __shfl_up_sync(0xffffffffu, ty[step], 1), tyr = __shfl_down_sync(0xffffffffu, ty[step], 1);
                tx[step] = txl - txr;                                      // hx of row Rb + 1 + step
                ty[step] = fmaf(2.f, ty[step], tyl + tyr);                 // vy
            }
#pragma unroll
            for (int step = 0; step < kRB; ++step) {
                const float hx = tx[step], vy = ty[step];
                if (s_own) s_gdst[step * (kTMC + 4)] = (s_hxA + s_vyA) + fmaf(2.f, s_hxB, hx - vy) + pg[step];
                s_hxA = s_hxB; s_hxB = hx; s_vyA = s_vyB; s_vyB = vy;
            }
        } else if (!EXT && p.do_sobel) {
#pragma unroll 2
            for (int step = 0; step < kRB; ++step) {
                const int qp = Rb + 2 + step;                       // input row q'
                int rr = (qp < 0) ? -qp : ((qp >= p.H) ? 2 * p.H - 2 - qp : qp);
                const int lr = min(max(rr - R0, b * kRB), b * kRB + 3 * kRB - 1) & (SmemB::kRows - 1);
                const float2 um = f2(sm.ring[0][lr][s_tm], sm.ring[1][lr][s_tm]);
                const float2 uc = f2(sm.ring[0][lr][s_t0], sm.ring[1][lr][s_t0]);
                const float2 up = f2(sm.ring[0][lr][s_tp], sm.ring[1][lr][s_tp]);
                const float umy = sm.ring[2][lr][s_tm], ucy = sm.ring[2][lr][s_t0], upy = sm.ring[2][lr][s_tp];
                const float2 d = fma2(bcast(-1.f), um, up);
                const float2 sv = fma2(bcast(2.f), uc, add2(um, up));
                const float2 gx = fma2(bcast(2.f), s_dB, add2(s_dA, d));   // Sobel of row q'-1
                const float2 gy = fma2(bcast(-1.f), s_sA, sv);
                const float dy = upy - umy;
                const float sy = fmaf(2.f, ucy, umy + upy);
                const float gxy = fmaf(2.f, s_dBy, s_dAy + dy);
                const float gyy = sy - s_sAy;
                const int qt = qp - 1;
                float tx = 0.f, ty = 0.f;
                if (s_colok && qt >= 0 && qt < p.H) {
                    const float S1 = fabsf(gx.x) + fabsf(gy.x), S2 = fabsf(gx.y) + fabsf(gy.y), Sy = fabsf(gxy) + fabsf(gyy);
                    float r;
                    if (FAST) r = mulsign(k_grad, Sy - fmaxf(S1, S2));
                    else if (p.grad_combine == MMIF_COMBINE_MAX) r = k_grad * norm_der(Sy - fmaxf(S1, S2), p.grad_norm);
                    else r = k_grad * (norm_der(Sy - S1, p.grad_norm) + norm_der(Sy - S2, p.grad_norm));
                    tx = mulsign(r, gxy);
                    ty = mulsign(r, gyy);
                    if (ZMODE && s_own && qt >= i0 && qt < iend) {
                        if (FAST || p.grad_combine == MMIF_COMBINE_MAX) z_grad += norm_val(Sy - fmaxf(S1, S2), FAST ? MMIF_NORM_L1 : p.grad_norm);
                        else z_grad += 0.5f * (norm_val(Sy - S1, p.grad_norm) + norm_val(Sy - S2, p.grad_norm));
                    }
                }
                const float txl = __shfl_up_sync(0xffffffffu, tx, 1), txr = __shfl_down_sync(0xffffffffu, tx, 1);
                const float tyl = __shfl_up_sync(0xffffffffu, ty, 1), tyr = __shfl_down_sync(0xffffffffu, ty, 1);
                float hx = txl - txr, vy = tyl + 2.f * ty + tyr;
                if (s_c == 1) { hx -= txl; vy += tyl; }                 // column -1 folds onto column 1
                if (s_c == p.W - 2) { hx += txr; vy += tyr; }           // column W folds onto column W-2
                const int gi = qp - 2;                                  // gradient row completed by this step
                float G = s_hxA + 2.f * s_hxB + hx + s_vyA - vy;
                if (gi == 1) G += s_hxA - s_vyA;                        // row -1 folds onto row 1
                if (gi == p.H - 2) G += hx + vy;                        // row H folds onto row H-2
                if (FAST) G += mulsign(k_pix, s_ucAy - fmaxf(s_ucA.x, s_ucA.y));
                else if (p.pixel_combine == MMIF_COMBINE_MAX) G += k_pix * norm_der(s_ucAy - fmaxf(s_ucA.x, s_ucA.y), p.pixel_norm);
                else G += k_pix * (norm_der(s_ucAy - s_ucA.x, p.pixel_norm) + norm_der(s_ucAy - s_ucA.y, p.pixel_norm));
                if (s_own && gi >= i0 && gi < iend) {
                    sb.gbuf[step][s_ci] = G;
                    if (ZMODE) {
                        if (FAST || p.pixel_combine == MMIF_COMBINE_MAX) z_pix += norm_val(s_ucAy - fmaxf(s_ucA.x, s_ucA.y), FAST ? MMIF_NORM_L1 : p.pixel_norm);
                        else z_pix += 0.5f * (norm_val(s_ucAy - s_ucA.x, p.pixel_norm) + norm_val(s_ucAy - s_ucA.y, p.pixel_norm));
                    }
                }
                s_dA = s_dB; s_dB = d; s_sA = s_sB; s_sB = sv; s_dAy = s_dBy; s_dBy = dy; s_sAy = s_sBy; s_sBy = sy;
                s_hxA = s_hxB; s_hxB = hx; s_vyA = s_vyB; s_vyB = vy;
                s_ucA = s_ucB; s_ucB = uc; s_ucAy = s_ucBy; s_ucBy = ucy;
            }
        }
        // ---------------- A1: vertical moments of window rows [Rb, Rb+8) ------------------------
        vpass_moments<WIN>(sm, p.taps, sh, (b & 3) * kRB, kVOFF);
        __syncthreads();
        // ---------------- A2: horizontal moments -> derivative coefficients -> cbuf -------------
        {
            const int q = Rb + ho;                     // window row
            float2 ab[8], cc[8];
            const bool active = (hg * 8 < kWC) && (q >= 0) && (q < p.Hout);
            if (active) {
                float2 acc[8][4];
                hpass<WIN, 4, false>(sm.vbuf + ho * kVPitch + hg * 8, kVCols, p.taps, acc);
                const unsigned zm = (ZMODE && q >= i0 && q < iend) ? a_zmask : 0u;
                // Branch-free over the 8 window columns (independent chains interleave); columns without a
                // valid window are computed on zero-filled data (finite) and masked at the end.
                // Stored coefficients: ab = (a/2, -b), cc = c/2 with a = dS/dmu_y', b = dS/dE[y'^2], c = dS/dE[x'y'].
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const Moments mo = moments_of(acc[j]);
                    const Stats st = stats_from(mo, sh);
                    const float myk = (st.vy >= 0.f) ? 1.f : 0.f;        // clamp(min=0) passes the gradient at 0
                    const float2 vk = max2(st.vk, 0.f);
                    const float vy = fmaxf(st.vy, 0.f);
                    const float2 A1 = fma2(st.mu, bcast(2.f * st.muy), bcast(p.C1));
                    const float2 B1 = fma2(st.mu, st.mu, bcast(fmaf(st.muy, st.muy, p.C1)));
                    const float2 A2 = fma2(bcast(2.f), st.cov, bcast(p.C2));
                    const float2 B2 = add2(vk, bcast(vy + p.C2));
                    const float2 R1 = rcp2(B1), R2 = rcp2(B2);
                    const float2 Cs = mul2(A2, R2);
                    float2 S, ch, nb, a;
                    if (EXT && p.cs_only) {                                          // S = cs = A2 / B2
                        S = Cs;
                        ch = R2;
                        nb = muls(myk, mul2(S, R2));
                        a = f2(0.f, 0.f);
                    } else {
                        const float2 L = mul2(A1, R1);
                        S = mul2(L, Cs);
                        ch = mul2(L, R2);                                            // (dS/dcov) / 2 = A1 / (B1 B2)
                        nb = muls(myk, mul2(S, R2));                                 // -dS/dvar_y = S / B2
                        a = mul2(mul2(Cs, R1), fma2(L, bcast(-st.muy), st.mu));      // (dS/dmu_y) / 2 = Cs / B1 (mu_k - mu_y L)
                    }
                    // a/2 = (dS/dmu_y)/2 - dS/dvar_y (my' + eps cy) - (dS/dcov)/2 (mk' + eps ck)
                    a = fma2(nb, bcast(mo.my + sh.ecy), a);
                    a = fma2(ch, fma2(mo.mk, bcast(-1.f), sh.nec), a);
                    const bool vj = (a_vmask >> j) & 1u;
                    if (EXT) {
                        float2 wq = pairw;
                        if (p.msw) {                                                    // gamma = sigma1 / (sigma1 + sigma2), loss.py:232-233
                            const float sg1 = fmaxf(vk.x, 1e-4f), sg2 = fmaxf(vk.y, 1e-4f);
                            const float gm = __fdiv_rn(sg1, fmaxf(sg1 + sg2, 1e-7f));
                            wq = f2(gm, 1.f - gm);
                        }
                        ab[j] = vj ? f2(wq.x * a.x + wq.y * a.y, wq.x * nb.x + wq.y * nb.y) : f2(0.f, 0.f);
                        cc[j] = vj ? mul2(ch, wq) : f2(0.f, 0.f);
                    } else {
                        ab[j] = vj ? f2(a.x + a.y, nb.x + nb.y) : f2(0.f, 0.f);
                        cc[j] = vj ? ch : f2(0.f, 0.f);
                    }
                    if (ZMODE) {
                        // SELECT, never multiply by 0: the two window columns past the strip read the never-written pad
                        // columns of vbuf, whose stale bits can be NaN / Inf (0 * NaN = NaN poisoned the loss sums)
                        const bool zq = (zm >> j) & 1u;
                        const float2 sg = max2(vk, 1e-4f);
                        z_ss = add2(z_ss, f2(zq ? S.x : 0.f, zq ? S.y : 0.f));
                        z_cs = add2(z_cs, f2(zq ? Cs.x : 0.f, zq ? Cs.y : 0.f));
                        z_sg = add2(z_sg, f2(zq ? sg.x : 0.f, zq ? sg.y : 0.f));
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
                // gradient column g sums window columns [g, g + HALO] of the tile (window 0 = image column j0 - HALO)
                hpass<WIN, 2, true>(tbuf + ho * kCPitch + hg * 8, kTWI, p.taps, acc);
                const int lr = (b * kRB + ho) & (SmemB::kRows - 1);
                float outv[8], u1[8], u2[8], uy[8], gb[8];
                {   // 8 pixels of row lr: LDS.128, lanes 0-7 are 8 ring rows (pitch = 16 mod 128 B)
                    const int tc = kOFF + hg * 8;
                    const float4* q1 = reinterpret_cast<const float4*>(&sm.ring[0][lr][tc]);
                    const float4* q2 = reinterpret_cast<const float4*>(&sm.ring[1][lr][tc]);
                    const float4* qy = reinterpret_cast<const float4*>(&sm.ring[2][lr][tc]);
                    const float4* qg = reinterpret_cast<const float4*>(&sb.gbuf[ho][hg * 8]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 a = q1[h], bq = q2[h], c = qy[h], g4 = qg[h];
                        u1[4 * h] = a.x; u1[4 * h + 1] = a.y; u1[4 * h + 2] = a.z; u1[4 * h + 3] = a.w;
                        u2[4 * h] = bq.x; u2[4 * h + 1] = bq.y; u2[4 * h + 2] = bq.z; u2[4 * h + 3] = bq.w;
                        uy[4 * h] = c.x; uy[4 * h + 1] = c.y; uy[4 * h + 2] = c.z; uy[4 * h + 3] = c.w;
                        gb[4 * h] = g4.x; gb[4 * h + 1] = g4.y; gb[4 * h + 2] = g4.z; gb[4 * h + 3] = g4.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x1s = u1[j] - sh.c.x;
                    const float x2s = u2[j] - sh.c.y;
                    const float ys = uy[j] - sh.cy;
                    const float dS = fmaf(x2s, acc[j][1].y, fmaf(x1s, acc[j][1].x, fmaf(-ys, acc[j][0].y, acc[j][0].x)));   // dS / 2
                    outv[j] = fmaf(k_ssim2, dS, gb[j]);
                }
                float* dst = p.dF + img_off + (size_t)i * p.W + j0 + hg * 8;
                if (EXT && p.accum) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j0 + hg * 8 + j < jend) outv[j] += dst[j];
                }
                if (p.vec_store && j0 + hg * 8 + 8 <= jend) {
                    reinterpret_cast<float4*>(dst)[0] = make_float4(outv[0], outv[1], outv[2], outv[3]);
                    reinterpret_cast<float4*>(dst)[1] = make_float4(outv[4], outv[5], outv[6], outv[7]);
                } else if (p.vec_store && j0 + hg * 8 + 4 == jend) {      // half group at the strip end (TG = 4 mod 8) / image edge
                    reinterpret_cast<float4*>(dst)[0] = make_float4(outv[0], outv[1], outv[2], outv[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j0 + hg * 8 + j < jend) dst[j] = outv[j];
                }
            }
        }
        __syncthreads();
    }
    if (ZMODE) {
        double v[8] = {(double)z_ss.x, (double)z_ss.y, (double)z_cs.x, (double)z_cs.y, (double)z_sg.x, (double)z_sg.y,
                       (double)z_pix, (double)z_grad};
        cta_finish<kNT>(p.fin, v, sm.red, &sm.flag, n, seg * p.nstrip + strip, p.nstrip * p.nseg);
    }
}

}  // namespace mmif
#include "fusion_loss_ws.cuh"
namespace mmif {

// Companion of the early exit above: dF = g * dF_unit when the three upstream gradients are equal,
// nothing otherwise (the recomputing kernel then does the work).  Pure streaming, 8 B/pixel.
__global__ void __launch_bounds__(256)
rescale_unit_kernel(const float* __restrict__ gs, const float* __restrict__ gp, const float* __restrict__ gg,
                    const float* __restrict__ unit, float* __restrict__ dF, size_t n, int vec) {
    const float g0 = gs ? __ldg(gs) : 0.f, g1 = gp ? __ldg(gp) : 0.f, g2 = gg ? __ldg(gg) : 0.f;
    if (!(g0 == g1 && g1 == g2)) return;
    if (unit == dF && g0 == 1.f) return;          // in place with unit upstream (total.backward()): the buffer already is dL/dIf
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const float4* u4 = reinterpret_cast<const float4*>(unit);
        float4* d4 = reinterpret_cast<float4*>(dF);
        const size_t n4 = n >> 2;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __ldcs(u4 + i + k * stride);
#pragma unroll
            for (int k = 0; k < 4; ++k) __stcs(d4 + i + k * stride, make_float4(g0 * v[k].x, g0 * v[k].y, g0 * v[k].z, g0 * v[k].w));
        }
        for (; i < n4; i += stride) {
            const float4 v = __ldcs(u4 + i);
            __stcs(d4 + i, make_float4(g0 * v.x, g0 * v.y, g0 * v.z, g0 * v.w));
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dF[(n4 << 2) + threadIdx.x] = g0 * unit[(n4 << 2) + threadIdx.x];
    } else {
        for (; i < n; i += stride) dF[i] = g0 * __ldg(unit + i);
    }
}

// =============================================================================== host side
static int check_common(const void* a, const void* b, const void* c, int B, int H, int W) {
    if (!a || !b || !c) { set_error("null image pointer"); return MMIF_E_NULL; }
    if (B < 1 || H < WIN11 || W < WIN11) { set_error("shape (%d,%d,%d): need B>=1 and H,W >= %d", B, H, W, WIN11); return MMIF_E_SHAPE; }
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

// workspace = [counters + partials (max of the forward / backward grids)][B x 8 per-sample sums]
static size_t loss_ws_core_bytes(int B, int H, int W) {
    if (!fwd_ws_bytes(WIN11, B, H, W)) return 0;
    size_t best = 0;
    const int wins[5] = {11, 9, 7, 5, 3};          // MSW_SSIM runs the same kernels with the smaller windows
    for (int k = 0; k < 5; ++k) {
        const size_t f = fwd_ws_bytes(wins[k], B, H, W);
        const BwdGeom g = bwd_geom(B, H, W, wins[k]);
        const size_t z = ws_counters_bytes(B) + (size_t)B * g.nstrip * g.nseg * 8 * sizeof(double);
        best = best > f ? best : f;
        best = best > z ? best : z;
        if (wins[k] == WIN11) {
            const BwdGeom gw = bwd_geom(B, H, W, WIN11, true);
            const size_t zw = ws_counters_bytes(B) + (size_t)B * geom_ctas_per_sample(gw) * 8 * sizeof(double);
            best = best > zw ? best : zw;
        }
    }
    return best;
}
extern "C" size_t mmif_loss_workspace_bytes(int B, int H, int W) {
    const size_t f = loss_ws_core_bytes(B, H, W);
    return f ? f + (size_t)B * 8 * sizeof(double) : 0;
}
extern "C" int mmif_loss_geometry(int B, int H, int W, int kernel, int* out10) {
    if (!out10) { set_error("null out"); return MMIF_E_NULL; }
    if (B < 1 || H < WIN11 || W < WIN11) { set_error("shape (%d,%d,%d): need B>=1 and H,W >= %d", B, H, W, WIN11); return MMIF_E_SHAPE; }
    const BwdGeom g = bwd_geom(B, H, W, WIN11, kernel != 0);
    const int v[10] = {g.nstrip, g.nseg, g.n_tall, g.seg_rows, g.seg_short, g.fine_strips, g.fine_rows, g.nseg_fine,
                       kernel != 0 ? geom_ctas_per_sample(g) : g.nstrip * g.nseg, bwd_tg(WIN11)};
    for (int i = 0; i < 10; ++i) out10[i] = v[i];
    return MMIF_OK;
}
extern "C" size_t mmif_loss_out_doubles(int B) { const size_t nd = loss_block_doubles(B); return nd + (nd + 1) / 2; }

struct BwdExtra { const float* pair_w; float ssim_base; int cs_only; int do_sobel; bool use_base; int win; double sigma; int msw; int accum; };

struct Upstream { const float* g[3]; };
static int launch_bwd(const float* i1, const float* i2, const float* f, int B, int H, int W, const MmifLossCfg* cfg,
                      const Upstream& up, const float* dF_unit, float* dF, int zmode, double* out, void* ws, size_t ws_bytes,
                      cudaStream_t st, const BwdExtra* ex = nullptr) {
    const int win = (ex && ex->win) ? ex->win : WIN11;
    // the training objective runs on the warp-specialised kernel; MMIF_LOSS_WS=0 is the A/B switch back to the 2-CTA kernel
    const char* ws_env = getenv("MMIF_LOSS_WS");            // read per call: the A/B tools flip it inside one process
    const bool ws_off = ws_env != nullptr && atoi(ws_env) == 0;
    const bool fast = cfg->pixel_combine == MMIF_COMBINE_MAX && cfg->grad_combine == MMIF_COMBINE_MAX &&
                      cfg->pixel_norm == MMIF_NORM_L1 && cfg->grad_norm == MMIF_NORM_L1;
    // (the other combine / norm modes run the general Sobel path with two norm derivatives per pixel on every batch; that makes
    // the Sobel group the longest pipeline stage: measured 22.6 vs 20.9 ms at 64x3072x4096, 2.79 vs 2.73 ms at 8x — although
    // 58 vs 65 us at 1x1024x1224 — so they stay on the 2-CTA kernel; MMIF_WS_ALL_MODES=1 puts them on the warp-specialised one)
    // ... at LARGE shapes; up to ~16 Mpix per launch the warp-specialised kernel wins there too (avg + l2: 36 vs 48 us at 2x512x512,
    // 0.491 vs 0.529 ms at 4x2048x2048), so those modes take it up to 32 Mpix.  MMIF_WS_ALL_MODES=1 forces it at any size.
    static const bool ws_all_modes = getenv("MMIF_WS_ALL_MODES") != nullptr;
    const bool small_enough = (long long)B * H * W <= (32ll << 20);
    const bool use_ws = !ex && win == WIN11 && (fast || ws_all_modes || small_enough) && !ws_off;
    const BwdGeom g = bwd_geom(B, H, W, win, use_ws);
    BwdParams p;
    memset(&p, 0, sizeof(p));
    p.x1 = i1; p.x2 = i2; p.y = f; p.dF = dF; p.dF_unit = dF_unit;
    p.gout[0] = up.g[0]; p.gout[1] = up.g[1]; p.gout[2] = up.g[2];
    p.B = B; p.H = H; p.W = W; p.Hout = g.Hout; p.Wout = g.Wout;
    p.seg_rows = g.seg_rows; p.seg_short = g.seg_short; p.n_tall = g.n_tall; p.nseg = g.nseg; p.nstrip = g.nstrip;
    p.fine_strips = g.fine_strips; p.fine_rows = g.fine_rows; p.nseg_fine = g.nseg_fine;
    make_taps(&p.taps, win, (ex && ex->win) ? ex->sigma : 1.5);
    const double L = cfg->data_range;
    p.C1 = (float)((0.01 * L) * (0.01 * L)); p.C2 = (float)((0.03 * L) * (0.03 * L));
    p.pixel_combine = cfg->pixel_combine; p.grad_combine = cfg->grad_combine;
    p.pixel_norm = cfg->pixel_norm; p.grad_norm = cfg->grad_norm;
    p.w_ssim = cfg->w_ssim; p.w_pixel = cfg->w_pixel; p.w_grad = cfg->w_grad;
    p.vec_store = ((W & 3) == 0) && ((((uintptr_t)dF) & 15) == 0);
    p.ssim_base = cfg->w_ssim * (-0.5f) / (float)B;
    p.do_sobel = 1;
    if (ex) {
        p.pair_w = ex->pair_w; p.cs_only = ex->cs_only; p.do_sobel = ex->do_sobel; p.msw = ex->msw; p.accum = ex->accum;
        if (ex->use_base) p.ssim_base = ex->ssim_base;
    }
    if (zmode) {
        const size_t core = loss_ws_core_bytes(B, H, W);
        if (!ws || ws_bytes < core + (size_t)B * 8 * sizeof(double)) { set_error("workspace too small"); return MMIF_E_WORKSPACE; }
        unsigned char* w8 = (unsigned char*)ws;
        p.fin.B = B; p.fin.H = H; p.fin.W = W; p.fin.Hout = g.Hout; p.fin.Wout = g.Wout;
        p.fin.finalize = FIN_LOSS;
        p.fin.w_ssim = cfg->w_ssim; p.fin.w_pixel = cfg->w_pixel; p.fin.w_grad = cfg->w_grad;
        p.fin.counters = (unsigned*)w8;
        p.fin.partial = (double*)(w8 + ws_counters_bytes(B));
        p.fin.sums = (double*)(w8 + core); p.fin.sums_stride = 8; p.fin.out = out;
    }
    CUtensorMap m1, m2, my;
    p.use_tma = make_tensor_map(&m1, i1, B, H, W, kRPB, kRB) && make_tensor_map(&m2, i2, B, H, W, kRPB, kRB) &&
                make_tensor_map(&my, f, B, H, W, kRPB, kRB);
    if (!p.use_tma) { memset(&m1, 0, sizeof(m1)); memset(&m2, 0, sizeof(m2)); memset(&my, 0, sizeof(my)); }
    static unsigned long long attr_done = 0ull;
    if (first_use_on_device(&attr_done)) {
        const int sz = (int)sizeof(SmemBwd);
        const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<11, true, false, false>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<11, false, false, false>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<11, true, true, false>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<11, false, true, false>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<11, false, false, true>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<9, false, false, true>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<7, false, false, true>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<5, false, false, true>, at, sz));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_bwd_kernel<3, false, false, true>, at, sz));
        const int szw = (int)sizeof(SmemWS);
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_ws_kernel<11, true, 0>, at, szw));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_ws_kernel<11, true, 1>, at, szw));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_ws_kernel<11, true, 2>, at, szw));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_ws_kernel<11, false, 0>, at, szw));
        MMIF_CUDA(cudaFuncSetAttribute(fusion_loss_ws_kernel<11, false, 1>, at, szw));
    }
    dim3 grid(g.nstrip, B, g.nseg);
    if (!zmode && dF_unit) {
        const size_t n = (size_t)B * H * W;
        const int vec = (((uintptr_t)dF | (uintptr_t)dF_unit) & 15) == 0;
        rescale_unit_kernel<<<148 * 8, 256, 0, st>>>(up.g[0], up.g[1], up.g[2], dF_unit, dF, n, vec);
        MMIF_CUDA(cudaGetLastError());
        count_launch(MMIF_CNT_RESCALE);
        // total = l1 + l2 + l3; total.backward(): autograd hands the SAME device scalar to all three outputs, so the host
        // already knows the upstream gradients are equal and the recomputing kernel (which would exit at once) is not launched
        if (up.g[0] && up.g[0] == up.g[1] && up.g[1] == up.g[2]) return MMIF_OK;
    }
    const size_t sm = sizeof(SmemBwd);
    if (ex) {                         // SSIM-only extended launches ('w-ssim', MS-SSIM levels, MSW-SSIM windows)
        if (zmode || p.do_sobel) { set_error("extended backward is SSIM-only"); return MMIF_E_MODE; }
        switch (win) {
            case 11: fusion_loss_bwd_kernel<11, false, false, true><<<grid, kNT, sm, st>>>(m1, m2, my, p); break;
            case 9: fusion_loss_bwd_kernel<9, false, false, true><<<grid, kNT, sm, st>>>(m1, m2, my, p); break;
            case 7: fusion_loss_bwd_kernel<7, false, false, true><<<grid, kNT, sm, st>>>(m1, m2, my, p); break;
            case 5: fusion_loss_bwd_kernel<5, false, false, true><<<grid, kNT, sm, st>>>(m1, m2, my, p); break;
            case 3: fusion_loss_bwd_kernel<3, false, false, true><<<grid, kNT, sm, st>>>(m1, m2, my, p); break;
            default: set_error("no backward kernel for window %d", win); return MMIF_E_MODE;
        }
    } else if (use_ws) {
        const size_t smw = sizeof(SmemWS);
        const dim3 gridw((unsigned)((size_t)B * geom_ctas_per_sample(g)));       // linear: class by class, segment by segment
        if (zmode == 2 && fast) fusion_loss_ws_kernel<11, true, 2><<<gridw, kWsNT, smw, st>>>(m1, m2, my, p);
        else if (zmode) {
            if (fast) fusion_loss_ws_kernel<11, true, 1><<<gridw, kWsNT, smw, st>>>(m1, m2, my, p);
            else fusion_loss_ws_kernel<11, false, 1><<<gridw, kWsNT, smw, st>>>(m1, m2, my, p);
        } else {
            if (fast) fusion_loss_ws_kernel<11, true, 0><<<gridw, kWsNT, smw, st>>>(m1, m2, my, p);
            else fusion_loss_ws_kernel<11, false, 0><<<gridw, kWsNT, smw, st>>>(m1, m2, my, p);
        }
    } else if (zmode) {
        if (fast) fusion_loss_bwd_kernel<11, true, true, false><<<grid, kNT, sm, st>>>(m1, m2, my, p);
        else fusion_loss_bwd_kernel<11, false, true, false><<<grid, kNT, sm, st>>>(m1, m2, my, p);
    } else {
        if (fast) fusion_loss_bwd_kernel<11, true, false, false><<<grid, kNT, sm, st>>>(m1, m2, my, p);
        else fusion_loss_bwd_kernel<11, false, false, false><<<grid, kNT, sm, st>>>(m1, m2, my, p);
    }
    MMIF_CUDA(cudaGetLastError());
    count_launch(ex ? MMIF_CNT_SSIM_BWD_EXT : (zmode ? MMIF_CNT_LOSS_SINGLE_PASS : MMIF_CNT_LOSS_BWD));
    return MMIF_OK;
}

extern "C" int mmif_fusion_loss_fwd(const float* i1, const float* i2, const float* f, int B, int H, int W,
                                    const MmifLossCfg* cfg, double* out, float* dF_unit, void* ws, size_t ws_bytes,
                                    void* stream) {
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    rc = check_cfg(cfg);
    if (rc) return rc;
    if (!out) { set_error("null out"); return MMIF_E_NULL; }
    if (!ws || ws_bytes < mmif_loss_workspace_bytes(B, H, W)) { set_error("workspace too small"); return MMIF_E_WORKSPACE; }
    if (cfg->want_grad) {
        if (!dF_unit) { set_error("want_grad needs dF_unit"); return MMIF_E_NULL; }
        if (((uintptr_t)dF_unit) & 3) { set_error("dF_unit must be 4-byte aligned"); return MMIF_E_ALIGN; }
        return launch_bwd(i1, i2, f, B, H, W, cfg, Upstream{{nullptr, nullptr, nullptr}}, nullptr, dF_unit, cfg->want_grad == 2 ? 2 : 1, out,
                          ws, ws_bytes, (cudaStream_t)stream);
    }
    const size_t core = loss_ws_core_bytes(B, H, W);
    FwdLaunch L;
    memset(&L, 0, sizeof(L));
    L.win = WIN11; L.sigma = 1.5; L.epi = EPI_SSIM; L.finalize = FIN_LOSS; L.do_sobel = 1;
    L.data_range = cfg->data_range; L.cfg = *cfg;
    double* sums = (double*)((unsigned char*)ws + core);
    return launch_moment_fwd(L, i1, i2, f, B, H, W, sums, 8, out, ws, core, (cudaStream_t)stream);
}

extern "C" int mmif_fusion_loss_bwd(const float* i1, const float* i2, const float* f, int B, int H, int W,
                                    const MmifLossCfg* cfg, const float* gout3, const float* dF_unit, float* dF, void* ws,
                                    size_t ws_bytes, void* stream) {
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    rc = check_cfg(cfg);
    if (rc) return rc;
    if (!gout3 || !dF) { set_error("null gout3/dF"); return MMIF_E_NULL; }
    if ((((uintptr_t)dF) | ((uintptr_t)dF_unit)) & 3) { set_error("dF / dF_unit must be 4-byte aligned"); return MMIF_E_ALIGN; }
    return launch_bwd(i1, i2, f, B, H, W, cfg, Upstream{{gout3, gout3 + 1, gout3 + 2}}, dF_unit, dF, 0, nullptr, ws, ws_bytes,
                      (cudaStream_t)stream);
}

/* The same with the three upstream gradients as separate device scalars, as autograd hands them to the node
 * (a NULL pointer = that loss value received no gradient = 0; at least one must be given). */
extern "C" int mmif_fusion_loss_bwd3(const float* i1, const float* i2, const float* f, int B, int H, int W,
                                     const MmifLossCfg* cfg, const float* g_ssim, const float* g_pixel, const float* g_grad,
                                     const float* dF_unit, float* dF, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    rc = check_cfg(cfg);
    if (rc) return rc;
    if (!dF) { set_error("null dF"); return MMIF_E_NULL; }
    if (!g_ssim && !g_pixel && !g_grad) { set_error("no upstream gradient given"); return MMIF_E_NULL; }
    if ((((uintptr_t)dF) | ((uintptr_t)dF_unit) | (uintptr_t)g_ssim | (uintptr_t)g_pixel | (uintptr_t)g_grad) & 3) {
        set_error("dF / dF_unit / upstream gradients must be 4-byte aligned");
        return MMIF_E_ALIGN;
    }
    return launch_bwd(i1, i2, f, B, H, W, cfg, Upstream{{g_ssim, g_pixel, g_grad}}, dF_unit, dF, 0, nullptr, ws, ws_bytes,
                      (cudaStream_t)stream);
}

static double loss_sigma_of(int win) { return win == 11 ? 1.5 : 0.15 * (win - 1); }     // loss.py:34
static bool msw_win_ok(int win) { return win == 11 || win == 9 || win == 7 || win == 5 || win == 3; }

/* d/dIf of  scale * sum_n sum_k pair_w[n][k] * mean_windows(S_k or cs_k)(I_k[n], If[n])  times the device
 * scalar gout1[0]: the building block of 'w-ssim' (loss.py:259-266, per-sample gamma) and of the
 * MS-SSIM levels (loss.py:140-158: cs on levels 0..3, ssim on level 4, per-sample chain-rule factors). */
extern "C" int mmif_ssim_bwd_ex_win(const float* i1, const float* i2, const float* f, int B, int H, int W, int win, float data_range,
                                    const float* gout1, const float* pair_w, int cs_only, float scale, float* dF, void* ws,
                                    size_t ws_bytes, void* stream);

extern "C" int mmif_ssim_bwd_ex(const float* i1, const float* i2, const float* f, int B, int H, int W, float data_range,
                                const float* gout1, const float* pair_w, int cs_only, float scale, float* dF, void* ws,
                                size_t ws_bytes, void* stream) {
    return mmif_ssim_bwd_ex_win(i1, i2, f, B, H, W, WIN11, data_range, gout1, pair_w, cs_only, scale, dF, ws, ws_bytes, stream);
}

/* The same for SSIM(win_size) of the loss module with win = 11, 9, 7, 5 or 3 (window sigma by the loss rule, loss.py:34). */
extern "C" int mmif_ssim_bwd_ex_win(const float* i1, const float* i2, const float* f, int B, int H, int W, int win, float data_range,
                                    const float* gout1, const float* pair_w, int cs_only, float scale, float* dF, void* ws,
                                    size_t ws_bytes, void* stream) {
    if (!i1 || !i2 || !f) { set_error("null image pointer"); return MMIF_E_NULL; }
    if (!msw_win_ok(win)) { set_error("SSIM window %d unsupported (11, 9, 7, 5, 3)", win); return MMIF_E_MODE; }
    if (B < 1 || H < win || W < win) { set_error("shape (%d,%d,%d) smaller than the %d-tap window", B, H, W, win); return MMIF_E_SHAPE; }
    if (((uintptr_t)i1 | (uintptr_t)i2 | (uintptr_t)f) & 3) { set_error("image pointers must be 4-byte aligned"); return MMIF_E_ALIGN; }
    if (!gout1 || !dF) { set_error("null gout1/dF"); return MMIF_E_NULL; }
    if (((uintptr_t)dF | (uintptr_t)pair_w) & 3) { set_error("dF / pair_w must be 4-byte aligned"); return MMIF_E_ALIGN; }
    MmifLossCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.w_ssim = 1.f; cfg.data_range = data_range;
    cfg.pixel_combine = cfg.grad_combine = MMIF_COMBINE_MAX; cfg.pixel_norm = cfg.grad_norm = MMIF_NORM_L1;
    BwdExtra ex;
    memset(&ex, 0, sizeof(ex));
    ex.pair_w = pair_w; ex.ssim_base = scale; ex.cs_only = cs_only; ex.do_sobel = 0; ex.use_base = true;
    if (win != WIN11) { ex.win = win; ex.sigma = loss_sigma_of(win); }
    return launch_bwd(i1, i2, f, B, H, W, &cfg, Upstream{{gout1, nullptr, nullptr}}, nullptr, dF, 0, nullptr, ws, ws_bytes,
                      (cudaStream_t)stream, &ex);
}

/* calc_ssim(size_average=True) of the loss module for a window of 11, 9, 7, 5 or 3 taps (SSIM(win_size), loss.py:163-185;
 * sigma by the loss rule): the per-sample means ssim, cs, sigma of the pairs (i1, f) and (i2, f) in the loss block
 * layout (out[MMIF_LOSS_HEAD + 6 n ...]; the head holds 1 - mean ssim, 0, 0).  ws from mmif_loss_workspace_bytes. */
extern "C" int mmif_ssim_fwd_win(const float* i1, const float* i2, const float* f, int B, int H, int W, int win, float data_range,
                                 double* out, void* ws, size_t ws_bytes, void* stream) {
    if (!i1 || !i2 || !f || !out) { set_error("null pointer"); return MMIF_E_NULL; }
    if (!msw_win_ok(win)) { set_error("SSIM window %d unsupported (11, 9, 7, 5, 3)", win); return MMIF_E_MODE; }
    if (B < 1 || H < win || W < win) { set_error("shape (%d,%d,%d) smaller than the %d-tap window", B, H, W, win); return MMIF_E_SHAPE; }
    if (((uintptr_t)i1 | (uintptr_t)i2 | (uintptr_t)f) & 3) { set_error("image pointers must be 4-byte aligned"); return MMIF_E_ALIGN; }
    const size_t core = loss_ws_core_bytes(B, H, W);
    if (!core || !ws || ws_bytes < core + (size_t)B * 8 * sizeof(double)) { set_error("workspace too small"); return MMIF_E_WORKSPACE; }
    FwdLaunch L;
    memset(&L, 0, sizeof(L));
    L.win = win; L.sigma = loss_sigma_of(win); L.epi = EPI_SSIM; L.finalize = FIN_LOSS; L.do_sobel = 0;
    L.data_range = data_range;
    L.cfg.w_ssim = 1.f; L.cfg.data_range = data_range;
    L.cfg.pixel_combine = L.cfg.grad_combine = MMIF_COMBINE_MAX; L.cfg.pixel_norm = L.cfg.grad_norm = MMIF_NORM_L1;
    double* sums = (double*)((unsigned char*)ws + core);
    return launch_moment_fwd(L, i1, i2, f, B, H, W, sums, 8, out, ws, core, (cudaStream_t)stream);
}

/* size_average=False of calc_ssim (loss.py:52-110, metric.py:316-364): the SSIM / CS / clamped-variance MAPS of the
 * pairs (i1, f) and (i2, f), each [B][H-10][W-10] (any of the six may be NULL).  11-tap window, sigma 1.5. */
extern "C" int mmif_ssim_maps(const float* i1, const float* i2, const float* f, int B, int H, int W, float data_range,
                              float* ssim1, float* cs1, float* sigma1, float* ssim2, float* cs2, float* sigma2, void* ws,
                              size_t ws_bytes, void* stream) {
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    const size_t core = loss_ws_core_bytes(B, H, W);
    if (!ws || ws_bytes < core + (size_t)B * 8 * sizeof(double)) { set_error("workspace too small"); return MMIF_E_WORKSPACE; }
    FwdLaunch L;
    memset(&L, 0, sizeof(L));
    L.win = WIN11; L.sigma = 1.5; L.epi = EPI_MAPS; L.finalize = FIN_SUMS; L.data_range = data_range;
    L.cfg.pixel_norm = L.cfg.grad_norm = MMIF_NORM_L1;
    L.maps[0] = ssim1; L.maps[1] = cs1; L.maps[2] = sigma1; L.maps[3] = ssim2; L.maps[4] = cs2; L.maps[5] = sigma2;
    double* sums = (double*)((unsigned char*)ws + core);
    return launch_moment_fwd(L, i1, i2, f, B, H, W, sums, 8, nullptr, ws, core, (cudaStream_t)stream);
}

/* test.py:49-73 post-step in ONE pass over imgf: the per-sample SSIM of (i1, f) and (i2, f) (calc_ssim with data_range,
 * test.py:51-52) and the 8-bit image save_result / denorm write (common.py:74-81, data/transform.py:32-35). */
extern "C" int mmif_test_post(const float* i1, const float* i2, const float* f, int B, int H, int W, float data_range,
                              double* out, unsigned char* denorm_u8, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_common(i1, i2, f, B, H, W);
    if (rc) return rc;
    if (!out) { set_error("null out"); return MMIF_E_NULL; }
    if (!ws || ws_bytes < mmif_loss_workspace_bytes(B, H, W)) { set_error("workspace too small"); return MMIF_E_WORKSPACE; }
    const size_t core = loss_ws_core_bytes(B, H, W);
    FwdLaunch L;
    memset(&L, 0, sizeof(L));
    L.win = WIN11; L.sigma = 1.5; L.epi = EPI_SSIM; L.finalize = FIN_LOSS; L.do_sobel = 1; L.data_range = data_range;
    L.cfg.w_ssim = 1.f; L.cfg.w_pixel = 1.f; L.cfg.w_grad = 1.f; L.cfg.data_range = data_range;
    L.cfg.pixel_combine = L.cfg.grad_combine = MMIF_COMBINE_MAX; L.cfg.pixel_norm = L.cfg.grad_norm = MMIF_NORM_L1;
    L.denorm = denorm_u8;
    double* sums = (double*)((unsigned char*)ws + core);
    return launch_moment_fwd(L, i1, i2, f, B, H, W, sums, 8, out, ws, core, (cudaStream_t)stream);
}


/* One window size of MSW_SSIM.forward (loss.py:226-237): out_sums[n] = sum over window positions of
 * gamma*ssim(I1,If) + (1-gamma)*ssim(I2,If), gamma = sigma1/(sigma1+sigma2) per position (device doubles,
 * B of them, stride 8).  Window sigma follows the loss rule (loss.py:34). */
extern "C" int mmif_mswssim_fwd(const float* i1, const float* i2, const float* f, int B, int H, int W, int win, float data_range,
                                double* out_sums8, void* ws, size_t ws_bytes, void* stream) {
    if (!i1 || !i2 || !f || !out_sums8) { set_error("null pointer"); return MMIF_E_NULL; }
    if (!msw_win_ok(win)) { set_error("msw-ssim window %d unsupported (11, 9, 7, 5, 3)", win); return MMIF_E_MODE; }
    FwdLaunch L;
    memset(&L, 0, sizeof(L));
    L.win = win; L.sigma = loss_sigma_of(win); L.epi = EPI_MSW; L.finalize = FIN_SUMS; L.data_range = data_range;
    L.cfg.pixel_norm = L.cfg.grad_norm = MMIF_NORM_L1;
    return launch_moment_fwd(L, i1, i2, f, B, H, W, out_sums8, 8, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

/* Its backward: dF (+)= gout1[0] * scale * d(sum_n out_sums[n] / (Hout*Wout))/dIf. */
extern "C" int mmif_mswssim_bwd(const float* i1, const float* i2, const float* f, int B, int H, int W, int win, float data_range,
                                const float* gout1, float scale, int accumulate, float* dF, void* ws, size_t ws_bytes, void* stream) {
    if (!i1 || !i2 || !f || !gout1 || !dF) { set_error("null pointer"); return MMIF_E_NULL; }
    if (!msw_win_ok(win)) { set_error("msw-ssim window %d unsupported (11, 9, 7, 5, 3)", win); return MMIF_E_MODE; }
    if (B < 1 || H < win || W < win) { set_error("shape smaller than the window"); return MMIF_E_SHAPE; }
    if (((uintptr_t)i1 | (uintptr_t)i2 | (uintptr_t)f | (uintptr_t)dF) & 3) { set_error("pointers must be 4-byte aligned"); return MMIF_E_ALIGN; }
    MmifLossCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.w_ssim = 1.f; cfg.data_range = data_range;
    cfg.pixel_combine = cfg.grad_combine = MMIF_COMBINE_MAX; cfg.pixel_norm = cfg.grad_norm = MMIF_NORM_L1;
    BwdExtra ex;
    memset(&ex, 0, sizeof(ex));
    ex.ssim_base = scale; ex.use_base = true; ex.win = win; ex.sigma = loss_sigma_of(win); ex.msw = 1; ex.accum = accumulate;
    return launch_bwd(i1, i2, f, B, H, W, &cfg, Upstream{{gout1, nullptr, nullptr}}, nullptr, dF, 0, nullptr, ws, ws_bytes,
                      (cudaStream_t)stream, &ex);
}
