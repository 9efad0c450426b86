// Library plumbing: error text, device check, Gaussian taps, TMA tensor-map encoding.
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace mmif {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return MMIF_E_CUDA;
}

// Taps as the reference builds them (loss.py:24-30 / metric.py:290-296): python-double exp, rounded
// to float32, divided by their float32 sum.  The last bit of that float32 sum depends on the order
// torch's CPU reduction adds in (vector width of the host), and the window-sum deviation
// eps = sum(W2d) - 1 ~ 1e-7 that decides the sign of the variance of a flat region (clamped at 0 by
// the reference) flips with it.  Callers that have the reference's own table (the Python mirror
// computes it with the same torch ops) register it with mmif_set_gaussian_taps; otherwise the
// sequential-sum table below is used.
struct TapOverride { int win; double sigma; float w[kMaxWin]; };
static std::mutex g_tap_mu;
static TapOverride g_tap_tab[32];
static int g_tap_n = 0;

void make_taps(Taps* t, int win, double sigma) {
    float g[kMaxWin];
    bool found = false;
    {
        std::lock_guard<std::mutex> lk(g_tap_mu);
        for (int i = 0; i < g_tap_n; ++i)
            if (g_tap_tab[i].win == win && fabs(g_tap_tab[i].sigma - sigma) < 1e-12) {
                memcpy(g, g_tap_tab[i].w, sizeof(float) * win);
                found = true;
                break;
            }
    }
    if (!found) {
        const int c = win / 2;
        float sum = 0.f;
        for (int i = 0; i < win; ++i) {
            const double d = (double)(i - c);
            g[i] = (float)exp(-(d * d) / (2.0 * sigma * sigma));
        }
        for (int i = 0; i < win; ++i) sum += g[i];
        for (int i = 0; i < win; ++i) g[i] = g[i] / sum;
    }
    for (int i = 0; i < kMaxWin; ++i) t->w[i] = (i < win) ? g[i] : 0.f;
    double s2 = 0.0;   // sum of the reference's float32 2-D outer-product window (loss.py:36-37)
    for (int i = 0; i < win; ++i)
        for (int j = 0; j < win; ++j) s2 += (double)(float)(t->w[i] * t->w[j]);
    double s1 = 0.0;
    for (int i = 0; i < win; ++i) s1 += (double)t->w[i];
    t->wsum = (float)s2;
    t->weps = (float)(s2 - 1.0);
    t->wrho = (float)(s2 / (s1 * s1) - 1.0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        (void)cudaGetLastError();
    });
    return fn;
}

static std::atomic<unsigned long long> g_counts[MMIF_CNT_N];
void count_launch(int which, unsigned long long n) {
    if (which >= 0 && which < MMIF_CNT_N) g_counts[which].fetch_add(n, std::memory_order_relaxed);
}

// A tensor map is a pure function of (base, N, H, W, box): the encoding (a driver call of a few microseconds, three per
// launch) is memoised per thread — a training loop presents the same buffers every step.
struct MapMemo { const float* base; int N, H, W, bw, bh; CUtensorMap map; };
constexpr int kMapMemo = 32;
static thread_local MapMemo t_maps[kMapMemo];
static thread_local int t_map_n = 0, t_map_next = 0;

bool make_tensor_map(CUtensorMap* map, const float* base, int N, int H, int W, int box_w, int box_h) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    static const bool disabled = getenv("MMIF_NO_TMA") != nullptr;   // debugging aid: force the plain-load ring
    if (disabled) return false;
    if ((W & 3) != 0 || (((uintptr_t)base) & 15) != 0) return false;   // global strides must be 16-byte multiples
    for (int i = 0; i < t_map_n; ++i) {
        const MapMemo& m = t_maps[i];
        if (m.base == base && m.N == N && m.H == H && m.W == W && m.bw == box_w && m.bh == box_h) {
            memcpy(map, &m.map, sizeof(CUtensorMap));
            count_launch(MMIF_CNT_TMAP_HIT);
            return true;
        }
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t gstride[2] = {(cuuint64_t)W * 4ull, (cuuint64_t)W * (cuuint64_t)H * 4ull};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // cuTensorMapEncodeTiled is a DRIVER call and needs a current context; a thread that has not made a runtime call yet
    // (autograd's backward thread on its first launch) has none -> CUDA_ERROR_INVALID_CONTEXT and a silent fall-back to the
    // plain-load ring.  Bind the runtime's primary context of the current device to this thread first.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        int d = 0;
        if (cudaGetDevice(&d) == cudaSuccess && cudaSetDevice(d) == cudaSuccess) (void)cudaFree(nullptr);
        (void)cudaGetLastError();
        ctx_bound = true;
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {          // belt and braces: bind and retry once
        (void)cudaFree(nullptr);
        r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { count_launch(MMIF_CNT_TMAP_FAIL); return false; }
    count_launch(MMIF_CNT_TMAP_ENCODE);
    MapMemo& m = t_maps[t_map_next];
    m.base = base; m.N = N; m.H = H; m.W = W; m.bw = box_w; m.bh = box_h;
    memcpy(&m.map, map, sizeof(CUtensorMap));
    t_map_next = (t_map_next + 1) % kMapMemo;
    if (t_map_n < kMapMemo) ++t_map_n;
    return true;
}

}  // namespace mmif

extern "C" int mmif_set_gaussian_taps(int win, double sigma, const float* taps) {
    using namespace mmif;
    if (!taps) { set_error("null taps"); return MMIF_E_NULL; }
    if (win < 1 || win > kMaxWin) { set_error("window size %d out of range 1..%d", win, kMaxWin); return MMIF_E_SHAPE; }
    std::lock_guard<std::mutex> lk(g_tap_mu);
    int slot = -1;
    for (int i = 0; i < g_tap_n; ++i)
        if (g_tap_tab[i].win == win && fabs(g_tap_tab[i].sigma - sigma) < 1e-12) slot = i;
    if (slot < 0) {
        if (g_tap_n >= 32) { set_error("tap table full"); return MMIF_E_MODE; }
        slot = g_tap_n++;
    }
    g_tap_tab[slot].win = win;
    g_tap_tab[slot].sigma = sigma;
    memcpy(g_tap_tab[slot].w, taps, sizeof(float) * win);
    return MMIF_OK;
}
extern "C" int mmif_launch_counts(unsigned long long* out, int n) {
    if (!out) { mmif::set_error("null out"); return MMIF_E_NULL; }
    for (int i = 0; i < n; ++i) out[i] = (i < MMIF_CNT_N) ? mmif::g_counts[i].load(std::memory_order_relaxed) : 0ull;
    return MMIF_OK;
}
extern "C" int mmif_version(void) { return MMIF_VERSION; }
extern "C" const char* mmif_last_error(void) { return mmif::g_err; }
extern "C" int mmif_check_device(int dev) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { mmif::cuda_fail(e, "cudaGetDeviceProperties"); return MMIF_E_DEVICE; }
    if (p.major != 10) { mmif::set_error("device %d is sm_%d%d; libmmif_b200 needs sm_100", dev, p.major, p.minor); return MMIF_E_DEVICE; }
    return MMIF_OK;
}
