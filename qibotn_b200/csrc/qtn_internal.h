// Internal helpers shared by the host and device translation units.
#pragma once
#include "../../include/qibotn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
// Records `msg` as the thread-local last error and returns `code`.
int qtn_fail(int code, const char *msg);
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
#define QTN_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) return qtn_fail(QTN_ERR_CUDA, cudaGetErrorString(_e));   \
  } while (0)
#endif
