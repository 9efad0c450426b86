// Instantiations and host launcher of the forward strip-streaming kernel.
#include <math.h>
#include "moment_fwd.cuh"

namespace mmif {

// Rows per segment (a multiple of 8) for `rows` rows split among CTAs that each pay `extra_rows` rows of
// halo / prologue work, with `other_ctas` strips x samples and `slots` CTAs resident on the chip:
// minimise  waves x (seg_rows + extra_rows)  where waves counts the quantised tail of the last wave.
// `tail` > 0 (backward kernels, measured): a last wave that fills at most half of the slots costs only `tail` of a full
// one, because a CTA alone on its SM runs 1.39x faster than two sharing it; with it the model tracks a scan of forced
// segment heights (64 .. 3072 rows, B = 8 and 64 of 3072x4096) to +-1 %.  tail == 0: the older, more pessimistic tail term.
int pick_seg_rows(int rows, int other_ctas, int slots, int extra_rows, double tail) {
    int best_seg = 8;
    double best = 1e300;
    const int max_nseg = ceil_div(rows, 8);
    for (int nseg = 1; nseg <= max_nseg; ++nseg) {
        const int seg = ceil_div(ceil_div(rows, nseg), 8) * 8;
        if (ceil_div(rows, seg) != nseg) continue;
        const double w = (double)other_ctas * nseg / slots;
        const double frac = w - floor(w);
        const double waves = tail > 0.0 ? floor(w) + (frac > 1e-9 ? (frac <= 0.5 ? tail : 1.0) : 0.0) : fmax(ceil(w), w + 0.5);
        const double cost = waves * (seg + extra_rows);
        if (cost < best) { best = cost; best_seg = seg; }
        if (seg <= 16) break;
    }
    return best_seg;
}
int fwd_seg_rows(int rows, int other_ctas) { return pick_seg_rows(rows, other_ctas, 3 * 148, 14); }

static int two_of(int win) { return ((kTWI - (win - 1)) / 4) * 4; }

struct Geo { int Hout, Wout, nstrip, seg_rows, nseg; };
static Geo geo_of(int win, int B, int H, int W) {
    Geo g;
    g.Hout = H - (win - 1); g.Wout = W - (win - 1);
    g.nstrip = ceil_div(g.Wout, two_of(win));
    g.seg_rows = fwd_seg_rows(g.Hout, B * g.nstrip);
    g.nseg = ceil_div(g.Hout, g.seg_rows);
    return g;
}

size_t fwd_ws_bytes(int win, int B, int H, int W) {
    if (B < 1 || H < win || W < win) return 0;
    const Geo g = geo_of(win, B, H, W);
    return ws_counters_bytes(B) + (size_t)B * g.nstrip * g.nseg * 8 * sizeof(double);
}

template <int WIN, int EPI>
static int launch_t(const CUtensorMap& m1, const CUtensorMap& m2, const CUtensorMap& my, const FwdParams& p, dim3 grid,
                    cudaStream_t st) {
    static unsigned long long attr_done = 0ull;
    if (first_use_on_device(&attr_done)) {
        MMIF_CUDA(cudaFuncSetAttribute(moment_fwd_kernel<WIN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemF)));
    }
    moment_fwd_kernel<WIN, EPI><<<grid, kNT, sizeof(SmemF), st>>>(m1, m2, my, p);
    MMIF_CUDA(cudaGetLastError());
    count_launch((EPI == EPI_SSIM && WIN == 11 && p.do_sobel && p.fin.finalize == FIN_LOSS && !p.denorm) ? MMIF_CNT_LOSS_FWD
                                                                                                         : MMIF_CNT_MOMENT_FWD);
    return MMIF_OK;
}

int launch_moment_fwd(const FwdLaunch& L, const float* x1, const float* x2, const float* y, int B, int H, int W,
                      double* sums, long long sums_stride, double* out, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = fwd_ws_bytes(L.win, B, H, W);
    if (need == 0) { set_error("shape (%d,%d,%d) smaller than the %d-tap window", B, H, W, L.win); return MMIF_E_SHAPE; }
    if (!ws || ws_bytes < need) { set_error("workspace too small: %zu < %zu", ws_bytes, need); return MMIF_E_WORKSPACE; }
    const Geo g = geo_of(L.win, B, H, W);
    FwdParams p;
    memset(&p, 0, sizeof(p));
    p.x1 = x1; p.x2 = x2; p.y = y;
    p.B = B; p.H = H; p.W = W; p.Hout = g.Hout; p.Wout = g.Wout;
    p.seg_rows = g.seg_rows; p.nseg = g.nseg; p.nstrip = g.nstrip;
    make_taps(&p.taps, L.win, L.sigma);
    const double R = L.data_range;
    p.C1 = (float)((0.01 * R) * (0.01 * R)); p.C2 = (float)((0.03 * R) * (0.03 * R));
    p.pixel_combine = L.cfg.pixel_combine; p.grad_combine = L.cfg.grad_combine;
    p.pixel_norm = L.cfg.pixel_norm; p.grad_norm = L.cfg.grad_norm;
    p.w_ssim = L.cfg.w_ssim; p.w_pixel = L.cfg.w_pixel; p.w_grad = L.cfg.w_grad;
    p.do_sobel = L.do_sobel;
    for (int i = 0; i < 6; ++i) p.maps[i] = L.maps[i];
    p.denorm = L.denorm;
    p.plain_moments = L.plain_moments;
    unsigned char* w8 = (unsigned char*)ws;
    p.fin.B = B; p.fin.H = H; p.fin.W = W; p.fin.Hout = g.Hout; p.fin.Wout = g.Wout;
    p.fin.finalize = L.finalize;
    p.fin.w_ssim = L.cfg.w_ssim; p.fin.w_pixel = L.cfg.w_pixel; p.fin.w_grad = L.cfg.w_grad;
    p.fin.counters = (unsigned*)w8;
    p.fin.partial = (double*)(w8 + ws_counters_bytes(B));
    p.fin.sums = sums; p.fin.sums_stride = sums_stride; p.fin.out = out;
    CUtensorMap m1, m2, my;
    p.use_tma = make_tensor_map(&m1, x1, B, H, W, kTWI, kRB) && make_tensor_map(&m2, x2, B, H, W, kTWI, kRB) &&
                make_tensor_map(&my, y, B, H, W, kTWI, kRB);
    if (!p.use_tma) { memset(&m1, 0, sizeof(m1)); memset(&m2, 0, sizeof(m2)); memset(&my, 0, sizeof(my)); }
    dim3 grid(g.nstrip, g.nseg, B);
    if (L.epi == EPI_SSIM && L.win == 11) return launch_t<11, EPI_SSIM>(m1, m2, my, p, grid, st);
    if (L.epi == EPI_SSIM && !L.do_sobel) {           // SSIM(win_size = 9, 7, 5, 3) of the loss module (loss.py:163-185)
        switch (L.win) {
            case 9: return launch_t<9, EPI_SSIM>(m1, m2, my, p, grid, st);
            case 7: return launch_t<7, EPI_SSIM>(m1, m2, my, p, grid, st);
            case 5: return launch_t<5, EPI_SSIM>(m1, m2, my, p, grid, st);
            case 3: return launch_t<3, EPI_SSIM>(m1, m2, my, p, grid, st);
        }
    }
    if (L.epi == EPI_MAPS && L.win == 11) return launch_t<11, EPI_MAPS>(m1, m2, my, p, grid, st);
    if (L.epi == EPI_VIF) {
        switch (L.win) {
            case 17: return launch_t<17, EPI_VIF>(m1, m2, my, p, grid, st);
            case 9: return launch_t<9, EPI_VIF>(m1, m2, my, p, grid, st);
            case 5: return launch_t<5, EPI_VIF>(m1, m2, my, p, grid, st);
            case 3: return launch_t<3, EPI_VIF>(m1, m2, my, p, grid, st);
        }
    }
    if (L.epi == EPI_MSW) {
        switch (L.win) {
            case 11: return launch_t<11, EPI_MSW>(m1, m2, my, p, grid, st);
            case 9: return launch_t<9, EPI_MSW>(m1, m2, my, p, grid, st);
            case 7: return launch_t<7, EPI_MSW>(m1, m2, my, p, grid, st);
            case 5: return launch_t<5, EPI_MSW>(m1, m2, my, p, grid, st);
            case 3: return launch_t<3, EPI_MSW>(m1, m2, my, p, grid, st);
        }
    }
    set_error("no kernel instantiated for window %d / epilogue %d", L.win, L.epi);
    return MMIF_E_MODE;
}

}  // namespace mmif
