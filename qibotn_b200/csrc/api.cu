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
int g_tc_enable = 1;
int g_tc_min_log2 = 18;
int g_tc_persistent = 1;
int g_tc_terms = 3;
int g_tc_accum_split = 1;
int g_tc_prepack = 1;
int g_tc_pair = 5;
int g_tc_stagger = 1;
int g_tc_fp16 = 1;
int g_svd_wide = 1;
int g_svd_group_mb = 0;
int g_svd_reg = 4;
}  // namespace

extern "C" int qtn_option_tc_enable(void) { return g_tc_enable; }
extern "C" int qtn_option_tc_min_log2(void) { return g_tc_min_log2; }
extern "C" int qtn_option_tc_persistent(void) { return g_tc_persistent; }
extern "C" int qtn_option_tc_terms(void) { return g_tc_terms; }
extern "C" int qtn_option_tc_accum_split(void) { return g_tc_accum_split; }
extern "C" int qtn_option_tc_prepack(void) { return g_tc_prepack; }
extern "C" int qtn_option_tc_pair(void) { return g_tc_pair; }
extern "C" int qtn_option_tc_stagger(void) { return g_tc_stagger; }
extern "C" int qtn_option_tc_fp16(void) { return g_tc_fp16; }
extern "C" int qtn_option_svd_wide(void) { return g_svd_wide; }
extern "C" int qtn_option_svd_group_mb(void) { return g_svd_group_mb; }
extern "C" int qtn_option_svd_reg(void) { return g_svd_reg; }

extern "C" int qtn_set_option(const char *name, int value) {
  if (!name) return qtn_fail(QTN_ERR_ARG, "qtn_set_option: null name");
  if (!std::strcmp(name, "tc_enable")) { g_tc_enable = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "tc_min_log2")) { g_tc_min_log2 = value; return QTN_OK; }
  if (!std::strcmp(name, "tc_persistent")) { g_tc_persistent = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "tc_terms")) { g_tc_terms = value >= 4 ? 4 : 3; return QTN_OK; }
  if (!std::strcmp(name, "tc_accum_split")) { g_tc_accum_split = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "tc_prepack")) { g_tc_prepack = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "tc_pair")) { g_tc_pair = value < 0 ? 0 : value; return QTN_OK; }
  if (!std::strcmp(name, "tc_stagger")) { g_tc_stagger = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "tc_fp16")) { g_tc_fp16 = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "svd_wide")) { g_svd_wide = value != 0; return QTN_OK; }
  if (!std::strcmp(name, "svd_group_mb")) { g_svd_group_mb = value < 0 ? 0 : value; return QTN_OK; }
  if (!std::strcmp(name, "svd_reg")) { g_svd_reg = value < 0 ? 0 : (value > 5 ? 5 : value); return QTN_OK; }
  return qtn_fail(QTN_ERR_ARG, "qtn_set_option: unknown option");
}

extern "C" int qtn_get_option(const char *name, int *value) {
  if (!name || !value) return qtn_fail(QTN_ERR_ARG, "qtn_get_option: null argument");
  if (!std::strcmp(name, "tc_enable")) { *value = g_tc_enable; return QTN_OK; }
  if (!std::strcmp(name, "tc_min_log2")) { *value = g_tc_min_log2; return QTN_OK; }
  if (!std::strcmp(name, "tc_persistent")) { *value = g_tc_persistent; return QTN_OK; }
  if (!std::strcmp(name, "tc_terms")) { *value = g_tc_terms; return QTN_OK; }
  if (!std::strcmp(name, "tc_accum_split")) { *value = g_tc_accum_split; return QTN_OK; }
  if (!std::strcmp(name, "tc_prepack")) { *value = g_tc_prepack; return QTN_OK; }
  if (!std::strcmp(name, "tc_pair")) { *value = g_tc_pair; return QTN_OK; }
  if (!std::strcmp(name, "tc_stagger")) { *value = g_tc_stagger; return QTN_OK; }
  if (!std::strcmp(name, "tc_fp16")) { *value = g_tc_fp16; return QTN_OK; }
  if (!std::strcmp(name, "svd_wide")) { *value = g_svd_wide; return QTN_OK; }
  if (!std::strcmp(name, "svd_group_mb")) { *value = g_svd_group_mb; return QTN_OK; }
  if (!std::strcmp(name, "svd_reg")) { *value = g_svd_reg; return QTN_OK; }
  return qtn_fail(QTN_ERR_ARG, "qtn_get_option: unknown option");
}

extern "C" int64_t qtn_sizeof(const char *name) {
  if (!name) return -1;
  if (!std::strcmp(name, "qtn_pair_desc")) return (int64_t)sizeof(qtn_pair_desc);
  if (!std::strcmp(name, "qtn_pair_plan_t")) return (int64_t)sizeof(qtn_pair_plan_t);
  if (!std::strcmp(name, "qtn_svd_item")) return (int64_t)sizeof(qtn_svd_item);
  if (!std::strcmp(name, "qtn_svd_trunc")) return (int64_t)sizeof(qtn_svd_trunc);
  if (!std::strcmp(name, "qtn_gemm_item")) return (int64_t)sizeof(qtn_gemm_item);
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
