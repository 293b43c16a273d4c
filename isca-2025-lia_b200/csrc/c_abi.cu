// Error plumbing, device info and small elementwise entry points of libliab200.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void lia_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int lia_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;   // B200
  }
  return sms;
}

bool lia_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LIA_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

extern "C" int lia_abi_version(void) { return LIA_ABI_VERSION; }
extern "C" const char* lia_last_error(void) { return g_err; }

extern "C" int lia_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  LIA_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  LIA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  LIA_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  LIA_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    lia_set_error("libliab200 is built for sm_100a only; device is sm_%d%d", maj, min);
    return LIA_ERR_ARCH;
  }
  return LIA_OK;
}

