// Library-wide plumbing: thread-local error text, device check, launch counter.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace sslam {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int query_device(int* sms) {
  int dev = 0, n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible (%s); libsslam_b200 has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return SSLAM_ENODEVICE;
  }
  SSLAM_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0;
  SSLAM_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SSLAM_CHECK_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10) {
    set_error("device %d has compute capability %d.x; libsslam_b200 is built for sm_100a only", dev,
              major);
    return SSLAM_ENODEVICE;
  }
  return SSLAM_OK;
}

static std::once_flag g_once;
static int g_dev_rc = SSLAM_ENODEVICE;
static int g_sms = 0;

int check_device() {
  std::call_once(g_once, [] { g_dev_rc = query_device(&g_sms); });
  if (g_dev_rc != SSLAM_OK && g_err[0] == 0)
    set_error("no usable sm_100 CUDA device; libsslam_b200 has no CPU fallback");
  return g_dev_rc;
}

int num_sms() { return g_sms > 0 ? g_sms : 148; }

}  // namespace sslam

extern "C" int sslam_abi_version(void) { return SSLAM_ABI_VERSION; }

extern "C" int sslam_last_error(char* buf, size_t len) {
  size_t n = strlen(sslam::g_err);
  if (buf && len) {
    size_t c = n < len - 1 ? n : len - 1;
    memcpy(buf, sslam::g_err, c);
    buf[c] = 0;
  }
  return (int)n;
}

extern "C" int sslam_device_check(void) { return sslam::check_device(); }

extern "C" uint64_t sslam_launch_count(void) { return sslam::g_launches.load(); }
