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
