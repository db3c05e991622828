// Shared host-side helpers for libtdgram.so (status codes, error text, segment plumbing).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "tdgram.h"

namespace td {

void set_error(const char* fmt, ...);

#define TD_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      td::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                    __LINE__);                                                           \
      return TD_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define TD_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      td::set_error(__VA_ARGS__);      \
      return TD_ERR_INVALID;           \
    }                                  \
  } while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
inline int64_t ceil_div(int64_t v, int64_t m) { return (v + m - 1) / m; }

// Validates one stream against the shared segment table.
int check_stream(const td_stream* s, const int64_t* seg_keep, int n_seg, const char* what);

inline int lags_of(const td_stream& s) { return s.pre + 1 + s.post; }
inline int64_t width_of(const td_stream& s) { return (int64_t)s.channels * lags_of(s); }

// Keep stream-ordered allocations cached in the device's default pool (the default release
// threshold of 0 hands the memory back to the driver at every synchronisation).
int ensure_mempool_config();

// td_shutdown helpers (gram_umma.cu, solve.cu).
void release_gram_events();
void release_solver_handles();

// Small host->device table upload into stream-ordered memory (freed with cudaFreeAsync).
template <typename T>
int upload(const std::vector<T>& host, T** dev, cudaStream_t st) {
  size_t bytes = host.size() * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  int rc = ensure_mempool_config();
  if (rc) return rc;
  TD_CHECK_CUDA(cudaMallocAsync((void**)dev, bytes, st));
  if (!host.empty())
    TD_CHECK_CUDA(cudaMemcpyAsync(*dev, host.data(), host.size() * sizeof(T),
                                  cudaMemcpyHostToDevice, st));
  return TD_OK;
}

// Device-side view of a stream (segment tables live in device memory).
struct StreamView {
  const float* data;
  int64_t ld;
  int channels, pre, lags;
  const int64_t* seg_start;
  const int64_t* seg_len;
};

}  // namespace td
