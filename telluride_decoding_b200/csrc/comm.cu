// Multi-GPU plumbing of libtdgram.so: NCCL communicator + the allreduce of the partial statistics.
//
// The reference has no multi-device path; what this replaces is the per-(lambda, fold) re-read of
// the data (regression.py:381-409): every rank accumulates the Gram of its share of the frames and
// ONE allreduce over NVLink combines them.  The Gram is symmetric, so only its upper
// triangle travels: td_allreduce_sym_f32 packs n(n+1)/2 floats (cfg 4: 1.87 GB instead of 3.75 GB),
// reduces them and unpacks into both triangles.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded if there is
// one, the system library otherwise), so libtdgram.so has no link-time dependency on it and a
// single-GPU user never touches it.  One process per GPU; the 128-byte unique id travels through
// whatever channel the host has (the Python layer uses its torch.distributed process group).
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace td {
namespace {

// The slice of nccl.h this file uses (stable NCCL 2 ABI).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

int nccl_api(NcclApi** out) {
  static NcclApi api;
  if (!api.lib) {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // torch's copy, if loaded
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
      set_error("td_comm: libnccl.so.2 not found (%s)", dlerror());
      return TD_ERR_UNSUPPORTED;
    }
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(lib, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(lib, "ncclGetVersion"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GetErrorString) {
      set_error("td_comm: libnccl.so.2 lacks an expected symbol");
      return TD_ERR_UNSUPPORTED;
    }
    api.lib = lib;
  }
  *out = &api;
  return TD_OK;
}

#define TD_CHECK_NCCL(api, expr)                                                           \
  do {                                                                                     \
    ncclResult_t r__ = (expr);                                                             \
    if (r__ != 0) {                                                                        \
      td::set_error("%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(r__), __FILE__, __LINE__); \
      return TD_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

// Row i of the upper triangle (columns i..n-1) starts at i*n - i(i-1)/2 in the packed vector.
__device__ __forceinline__ long long tri_row_start(long long i, long long n) { return i * n - i * (i - 1) / 2; }

__global__ void pack_upper_kernel(const float* __restrict__ G, long long n, long long ldg, float* __restrict__ packed) {
  for (long long i = blockIdx.x; i < n; i += gridDim.x) {
    const float* src = G + i * ldg;
    float* dst = packed + tri_row_start(i, n) - i;
    for (long long j = i + threadIdx.x; j < n; j += blockDim.x) dst[j] = src[j];
  }
}

__global__ void unpack_upper_kernel(const float* __restrict__ packed, long long n, long long ldg, float* __restrict__ G) {
  for (long long i = blockIdx.x; i < n; i += gridDim.x) {
    const float* src = packed + tri_row_start(i, n) - i;
    float* dst = G + i * ldg;
    for (long long j = i + threadIdx.x; j < n; j += blockDim.x) dst[j] = src[j];
  }
}

}  // namespace
}  // namespace td

using namespace td;

struct td_comm {
  void* nccl;      // ncclComm_t
  int rank, world;
};

extern "C" int td_comm_unique_id(void* id128) {
  TD_REQUIRE(id128 != nullptr, "td_comm_unique_id: null argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  ncclUniqueId id;
  TD_CHECK_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return TD_OK;
}

extern "C" int td_comm_init(td_comm** comm, const void* id128, int rank, int world) {
  TD_REQUIRE(comm && id128, "td_comm_init: null argument");
  TD_REQUIRE(world >= 1 && rank >= 0 && rank < world, "td_comm_init: rank %d of %d", rank, world);
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  TD_CHECK_NCCL(api, api->CommInitRank(&c, world, id, rank));
  *comm = new td_comm{c, rank, world};
  return TD_OK;
}

extern "C" int td_comm_destroy(td_comm* comm) {
  if (!comm) return TD_OK;
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  if (comm->nccl) TD_CHECK_NCCL(api, api->CommDestroy(static_cast<ncclComm_t>(comm->nccl)));
  delete comm;
  return TD_OK;
}

extern "C" int td_comm_nccl_version(int* version) {
  TD_REQUIRE(version != nullptr, "td_comm_nccl_version: null argument");
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  *version = 0;
  if (api->GetVersion) TD_CHECK_NCCL(api, api->GetVersion(version));
  return TD_OK;
}

extern "C" int td_allreduce_sum_f32(td_comm* comm, float* buf_dev, int64_t n, void* stream) {
  TD_REQUIRE(comm && (buf_dev || n == 0) && n >= 0, "td_allreduce_sum_f32: bad arguments");
  if (n == 0 || comm->world == 1) return TD_OK;
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  TD_CHECK_NCCL(api, api->AllReduce(buf_dev, buf_dev, (size_t)n, kNcclFloat32, kNcclSum,
                                    static_cast<ncclComm_t>(comm->nccl), (cudaStream_t)stream));
  return TD_OK;
}

extern "C" int td_allreduce_sum_f64(td_comm* comm, double* buf_dev, int64_t n, void* stream) {
  TD_REQUIRE(comm && (buf_dev || n == 0) && n >= 0, "td_allreduce_sum_f64: bad arguments");
  if (n == 0 || comm->world == 1) return TD_OK;
  NcclApi* api;
  int rc = nccl_api(&api);
  if (rc) return rc;
  TD_CHECK_NCCL(api, api->AllReduce(buf_dev, buf_dev, (size_t)n, kNcclFloat64, kNcclSum,
                                    static_cast<ncclComm_t>(comm->nccl), (cudaStream_t)stream));
  return TD_OK;
}

extern "C" int td_pack_upper_f32(const float* G_dev, int64_t n, int64_t ldg, float* packed_dev, void* stream) {
  TD_REQUIRE(G_dev && packed_dev && n >= 0 && ldg >= n, "td_pack_upper_f32: bad arguments");
  if (n == 0) return TD_OK;
  pack_upper_kernel<<<(unsigned)std::min<int64_t>(n, 148 * 16), 256, 0, (cudaStream_t)stream>>>(G_dev, n, ldg, packed_dev);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}

extern "C" int td_unpack_upper_f32(const float* packed_dev, int64_t n, int64_t ldg, float* G_dev, void* stream) {
  TD_REQUIRE(G_dev && packed_dev && n >= 0 && ldg >= n, "td_unpack_upper_f32: bad arguments");
  if (n == 0) return TD_OK;
  unpack_upper_kernel<<<(unsigned)std::min<int64_t>(n, 148 * 16), 256, 0, (cudaStream_t)stream>>>(packed_dev, n, ldg, G_dev);
  TD_CHECK_CUDA(cudaGetLastError());
  return td_symmetrize_f32(G_dev, n, ldg, /*from_upper=*/1, stream);
}

extern "C" int td_allreduce_sym_f32(td_comm* comm, float* G_dev, int64_t n, int64_t ldg, float* packed_dev,
                                    size_t packed_bytes, void* stream) {
  TD_REQUIRE(comm && G_dev && n >= 0 && ldg >= n, "td_allreduce_sym_f32: bad arguments");
  if (n == 0 || comm->world == 1) return TD_OK;
  const size_t need = (size_t)n * (size_t)(n + 1) / 2 * sizeof(float);
  if (!packed_dev || packed_bytes < need) {
    set_error("td_allreduce_sym_f32: packed buffer has %zu bytes, %zu needed", packed_bytes, need);
    return TD_ERR_WORKSPACE;
  }
  int rc = td_pack_upper_f32(G_dev, n, ldg, packed_dev, stream);
  if (rc) return rc;
  rc = td_allreduce_sum_f32(comm, packed_dev, (int64_t)(need / sizeof(float)), stream);
  if (rc) return rc;
  return td_unpack_upper_f32(packed_dev, n, ldg, G_dev, stream);
}
