// Internal helpers shared by the host and device translation units.
#pragma once
#include "../../include/qibotn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
// Records `msg` as the thread-local last error and returns `code`.
int qtn_fail(int code, const char *msg);
// The options a call runs with: a normalised copy of `given`, or of the process-wide set when NULL (api.cu).
void qtn_options_resolve(const qtn_options *given, qtn_options *out);
#ifdef __cplusplus
}
#endif

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define QTN_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) return qtn_fail(QTN_ERR_CUDA, cudaGetErrorString(_e));   \
  } while (0)
// gett_tc.cu: tcgen05 kernel launch (plan->kernel == QTN_KERNEL_TC)
// amax_*: device words bounding max(|re|, |im|) of each tensor (see qtn_pair_run_scaled); may be NULL
int qtn_launch_tc(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                  const int32_t *slice_id, const uint32_t *amax_a, const uint32_t *amax_b, uint32_t *amax_c,
                  cudaStream_t st);
// gett.cu: *amax = max(*amax, bits of max(|re|, |im|)) over n complex64 elements
int qtn_launch_absmax(const void *x, long long n, uint32_t *amax, cudaStream_t st);
// gett.cu: C (+)= sum of `nsplit` split-K partials of `n` elements each
int qtn_launch_splitk_sum(int dtype, void *c, const void *ws, long long n, int nsplit, int accumulate,
                          cudaStream_t st);
#define QTN_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) return qtn_fail(QTN_ERR_CUDA, cudaGetErrorString(_e));   \
  } while (0)
#endif
