// Status plumbing of libtdgram.so.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace td {
namespace {
thread_local char g_error[1024] = "";
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int ensure_mempool_config() {
  static bool done[64] = {false};
  int dev = 0;
  TD_CHECK_CUDA(cudaGetDevice(&dev));
  if (done[dev & 63]) return TD_OK;
  cudaMemPool_t pool;
  TD_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t threshold = UINT64_MAX;
  TD_CHECK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  done[dev & 63] = true;
  return TD_OK;
}
}  // namespace td

extern "C" const char* td_last_error(void) { return td::g_error; }
extern "C" int td_version(void) { return TD_VERSION; }
extern "C" int td_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  TD_CHECK_CUDA(cudaGetDevice(&dev));
  if (sm_count) TD_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) TD_CHECK_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) TD_CHECK_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return TD_OK;
}

// Releases what the library keeps between calls: the calling thread's timing events, the
// cuSOLVER / cuBLAS handles with their side streams and events.  Call once the caller's streams
// are idle; the library can be used again afterwards (everything is recreated on demand).
extern "C" int td_shutdown(void) {
  td::release_gram_events();
  td::release_solver_handles();
  return TD_OK;
}
