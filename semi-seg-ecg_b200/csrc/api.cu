// Library-level entry points of libsemiseg_b200: version, error reporting, device check.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

std::atomic<long long> g_ssb_launches{0};
int g_ssb_pdl = 2;   // env SSB_PDL: see common.cuh
int g_ssb_num_sms = 148;
static thread_local char g_err[512] = "";

void ssb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ssb_sm100_prepare();
int ssb_simt_prepare();
int ssb_loss_prepare();
int ssb_aug_prepare();

extern "C" {

int ssb_prepare(void) {
  const char* pdl = getenv("SSB_PDL");
  if (pdl) g_ssb_pdl = atoi(pdl);
  int rc = ssb_device_check();
  if (rc) return rc;
  {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      g_ssb_num_sms = sms;
  }
  if ((rc = ssb_simt_prepare())) return rc;
  if ((rc = ssb_loss_prepare())) return rc;
  if ((rc = ssb_aug_prepare())) return rc;
  return ssb_sm100_prepare();
}

int ssb_version(void) { return SSB_VERSION; }

const char* ssb_last_error(void) { return g_err; }

int64_t ssb_launch_count(void) { return (int64_t)g_ssb_launches.load(); }

int ssb_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    ssb_set_error("ssb_device_check: no CUDA device: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    ssb_set_error("ssb_device_check: device %d is sm_%d%d; this library is built for sm_100a only", dev,
                  major, minor);
    return SSB_ERR_UNSUPPORTED;
  }
  return SSB_OK;
}

int ssb_memset_zero(void* p, size_t bytes, ssb_stream_t stream) {
  if (bytes == 0) return SSB_OK;
  SSB_REQUIRE(p != nullptr, "ssb_memset_zero: null pointer");
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, to_stream(stream));
  if (e != cudaSuccess) {
    ssb_set_error("ssb_memset_zero: %s", cudaGetErrorString(e));
    return SSB_ERR_CUDA;
  }
  return SSB_OK;
}

}  // extern "C"
