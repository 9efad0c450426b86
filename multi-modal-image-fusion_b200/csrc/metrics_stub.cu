// Temporary: entry points not built yet return MMIF_E_MODE (never a silent fallback).
#include "common.cuh"
using namespace mmif;
#define NOT_BUILT(name) { set_error(name " is not built yet"); return MMIF_E_MODE; }
extern "C" int mmif_tv_loss(const float*, int, int, int, int, float, double*, void*, size_t, void*) NOT_BUILT("mmif_tv_loss")
extern "C" size_t mmif_metric_workspace_bytes(int, int, int) { return 256; }
extern "C" int mmif_stats(const float*, const float*, const float*, int, int, int, double*, void*, size_t, void*) NOT_BUILT("mmif_stats")
extern "C" int mmif_hist(const float*, const float*, const float*, int, int, int, uint32_t*, double*, void*, size_t, void*) NOT_BUILT("mmif_hist")
extern "C" int mmif_qabf(const float*, const float*, const float*, int, int, int, float, double*, void*, size_t, void*) NOT_BUILT("mmif_qabf")
extern "C" int mmif_ssim(const float*, const float*, const float*, int, int, int, int, float, int, double*, void*, size_t, void*) NOT_BUILT("mmif_ssim")
extern "C" int mmif_msssim(const float*, const float*, const float*, int, int, int, int, float, double*, void*, size_t, void*) NOT_BUILT("mmif_msssim")
extern "C" int mmif_viff(const float*, const float*, const float*, int, int, int, double*, void*, size_t, void*) NOT_BUILT("mmif_viff")
extern "C" int mmif_eval_suite(const float*, const float*, const float*, int, int, int, double*, void*, size_t, void*) NOT_BUILT("mmif_eval_suite")
extern "C" int mmif_eval_suite_host(const float*, const float*, const float*, int, int, int, double*, float*, double*, void*, size_t, void*) NOT_BUILT("mmif_eval_suite_host")
