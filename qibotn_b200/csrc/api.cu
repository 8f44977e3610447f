// Error reporting and device facts.
#include <cstring>
#include <string>

#include "qtn_internal.h"

namespace {
thread_local std::string g_last_error;
}

extern "C" int qtn_fail(int code, const char *msg) {
  g_last_error = msg ? msg : "";
  return code;
}

extern "C" const char *qtn_last_error(void) { return g_last_error.c_str(); }

namespace {
constexpr qtn_options kDefaults = {1, 18, 1, 3, 1, 1, 5, 1, 1, 1, 0, 4, 2, 1, 10, {0}};
qtn_options g_opt = kDefaults;

struct Field { const char *name; int32_t qtn_options::*member; };
const Field kFields[] = {
    {"tc_enable", &qtn_options::tc_enable},       {"tc_min_log2", &qtn_options::tc_min_log2},
    {"tc_persistent", &qtn_options::tc_persistent}, {"tc_terms", &qtn_options::tc_terms},
    {"tc_accum_split", &qtn_options::tc_accum_split}, {"tc_prepack", &qtn_options::tc_prepack},
    {"tc_pair", &qtn_options::tc_pair},           {"tc_stagger", &qtn_options::tc_stagger},
    {"tc_fp16", &qtn_options::tc_fp16},           {"svd_wide", &qtn_options::svd_wide},
    {"svd_group_mb", &qtn_options::svd_group_mb}, {"svd_reg", &qtn_options::svd_reg},
    {"tc_gather", &qtn_options::tc_gather},       {"tc_prepack_a", &qtn_options::tc_prepack_a},
    {"tc_gather_min_run", &qtn_options::tc_gather_min_run},
};

// the value ranges every option is clamped to, wherever it comes from
void normalise(qtn_options &o) {
  o.tc_enable = o.tc_enable != 0;
  o.tc_persistent = o.tc_persistent != 0;
  o.tc_terms = o.tc_terms >= 4 ? 4 : 3;
  o.tc_accum_split = o.tc_accum_split != 0;
  o.tc_prepack = o.tc_prepack != 0;
  o.tc_pair = o.tc_pair < 0 ? 0 : o.tc_pair;
  o.tc_stagger = o.tc_stagger != 0;
  o.tc_fp16 = o.tc_fp16 != 0;
  o.tc_gather = o.tc_gather < 0 ? 0 : (o.tc_gather > 2 ? 2 : o.tc_gather);
  o.tc_prepack_a = o.tc_prepack_a < 0 ? 0 : (o.tc_prepack_a > 2 ? 2 : o.tc_prepack_a);
  o.tc_gather_min_run = o.tc_gather_min_run < 6 ? 6 : (o.tc_gather_min_run > 11 ? 11 : o.tc_gather_min_run);
  o.svd_wide = o.svd_wide != 0;
  o.svd_group_mb = o.svd_group_mb < 0 ? 0 : o.svd_group_mb;
  o.svd_reg = o.svd_reg < 0 ? 0 : (o.svd_reg > 5 ? 5 : o.svd_reg);
}
}  // namespace

extern "C" void qtn_options_default(qtn_options *opt) { if (opt) *opt = kDefaults; }
extern "C" void qtn_options_current(qtn_options *opt) { if (opt) *opt = g_opt; }
// internal: the options a call runs with (a normalised copy of the caller's, or of the process-wide set)
extern "C" void qtn_options_resolve(const qtn_options *given, qtn_options *out) {
  *out = given ? *given : g_opt;
  normalise(*out);
}

extern "C" int qtn_set_option(const char *name, int value) {
  if (!name) return qtn_fail(QTN_ERR_ARG, "qtn_set_option: null name");
  for (const Field &f : kFields)
    if (!std::strcmp(name, f.name)) { g_opt.*(f.member) = value; normalise(g_opt); return QTN_OK; }
  return qtn_fail(QTN_ERR_ARG, "qtn_set_option: unknown option");
}

extern "C" int qtn_get_option(const char *name, int *value) {
  if (!name || !value) return qtn_fail(QTN_ERR_ARG, "qtn_get_option: null argument");
  for (const Field &f : kFields)
    if (!std::strcmp(name, f.name)) { *value = g_opt.*(f.member); return QTN_OK; }
  return qtn_fail(QTN_ERR_ARG, "qtn_get_option: unknown option");
}

extern "C" int64_t qtn_sizeof(const char *name) {
  if (!name) return -1;
  if (!std::strcmp(name, "qtn_pair_desc")) return (int64_t)sizeof(qtn_pair_desc);
  if (!std::strcmp(name, "qtn_pair_plan_t")) return (int64_t)sizeof(qtn_pair_plan_t);
  if (!std::strcmp(name, "qtn_svd_item")) return (int64_t)sizeof(qtn_svd_item);
  if (!std::strcmp(name, "qtn_svd_trunc")) return (int64_t)sizeof(qtn_svd_trunc);
  if (!std::strcmp(name, "qtn_gemm_item")) return (int64_t)sizeof(qtn_gemm_item);
  if (!std::strcmp(name, "qtn_options")) return (int64_t)sizeof(qtn_options);
  return -1;
}

extern "C" const char *qtn_version(void) { return "qibotn_b200 0.1.0 (sm_100a)"; }

extern "C" int qtn_device_info(int device, int *sm_count, int *max_smem_optin, int *cc_major,
                               int *cc_minor) {
  cudaDeviceProp prop;
  QTN_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (max_smem_optin) *max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return QTN_OK;
}


// ---------------------------------------------------------------------------------------------
// Measured FMA-pipe peak (bench.py's roofline denominator for the kernels that run on the FP32 / FP64
// FMA pipes: MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers).  Every thread runs 16
// independent FMA chains for `iters` rounds; returns 2 * FMAs / elapsed in TFLOP/s.
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) fma_peak_kernel(R *out, int iters, R b, R c) {
  R a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = (R)(threadIdx.x + j) * (R)1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = fma(a[j], b, c);
  }
  R s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += a[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int qtn_microbench_fma(int is_double, int iters, void *scratch, int64_t scratch_bytes, double *tflops,
                                  void *stream) {
  if (!scratch || !tflops || iters < 1) return qtn_fail(QTN_ERR_ARG, "qtn_microbench_fma: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  QTN_CUDA_CHECK(cudaGetDevice(&dev));
  QTN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256;
  if (scratch_bytes < (int64_t)blocks * threads * 8)
    return qtn_fail(QTN_ERR_WORKSPACE, "qtn_microbench_fma: scratch too small (148 * 8 * 256 * 8 bytes)");
  cudaEvent_t e0, e1;
  QTN_CUDA_CHECK(cudaEventCreate(&e0));
  QTN_CUDA_CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    QTN_CUDA_CHECK(cudaEventRecord(e0, st));
    if (is_double) fma_peak_kernel<double><<<blocks, threads, 0, st>>>((double *)scratch, iters, 0.999999, 1e-7);
    else fma_peak_kernel<float><<<blocks, threads, 0, st>>>((float *)scratch, iters, 0.999999f, 1e-7f);
    QTN_CUDA_CHECK(cudaEventRecord(e1, st));
    QTN_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    QTN_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  QTN_LAUNCH_CHECK();
  *tflops = 2.0 * 16.0 * (double)iters * (double)blocks * threads / ((double)best * 1e-3) / 1e12;
  return QTN_OK;
}
