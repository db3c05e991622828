// Ridge and CCA solves on the accumulated sufficient statistics.  cuSOLVER / cuBLAS in float64 --
// a stated non-hot-path dependency: the O(D^3) factorisations are tiny next to the O(T D^2) Gram.
//
//   td_ridge_solve : brain_model.py:447-454 (normalise, + lambda*I over every row, bias included)
//                    and :477-479 (np.linalg.solve, split weights / bias) for a whole lambda sweep
//   td_cca_solve   : cca.py:337-367 (covariances, whitening by eigendecomposition, SVD, rotations)
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace td {
namespace {

constexpr int MAX_LANES = 8;    // concurrent factorisations of a lambda sweep

struct Handles {
  cusolverDnHandle_t solver = nullptr;
  cublasHandle_t blas = nullptr;
  cublasHandle_t blas_side = nullptr;     // bound to lane_stream[0]: look-ahead trailing updates
  cudaEvent_t ev_panel[8] = {};
  cudaEvent_t ev_wide[8] = {};
  // side streams (+ their own cuSOLVER handles) so that the independent Cholesky factorisations of
  // a regularisation sweep overlap: one n ~ 2000 potrf cannot fill 148 SMs
  cudaStream_t lane_stream[MAX_LANES] = {};
  cusolverDnHandle_t lane_solver[MAX_LANES] = {};
  cudaEvent_t ev_ready = nullptr;
  cudaEvent_t ev_done[MAX_LANES] = {};
};

Handles g_handles[64];

int get_handles(Handles** out) {
  Handles* per_device = g_handles;
  int dev = 0;
  TD_CHECK_CUDA(cudaGetDevice(&dev));
  Handles& h = per_device[dev & 63];
  if (!h.solver) {
    if (cusolverDnCreate(&h.solver) != CUSOLVER_STATUS_SUCCESS) {
      set_error("cusolverDnCreate failed");
      return TD_ERR_SOLVER;
    }
    if (cublasCreate(&h.blas) != CUBLAS_STATUS_SUCCESS) {
      set_error("cublasCreate failed");
      return TD_ERR_SOLVER;
    }
    TD_CHECK_CUDA(cudaEventCreateWithFlags(&h.ev_ready, cudaEventDisableTiming));
    for (int i = 0; i < 8; ++i) {
      TD_CHECK_CUDA(cudaEventCreateWithFlags(&h.ev_panel[i], cudaEventDisableTiming));
      TD_CHECK_CUDA(cudaEventCreateWithFlags(&h.ev_wide[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < MAX_LANES; ++i) {
      TD_CHECK_CUDA(cudaStreamCreateWithFlags(&h.lane_stream[i], cudaStreamNonBlocking));
      TD_CHECK_CUDA(cudaEventCreateWithFlags(&h.ev_done[i], cudaEventDisableTiming));
      if (cusolverDnCreate(&h.lane_solver[i]) != CUSOLVER_STATUS_SUCCESS ||
          cusolverDnSetStream(h.lane_solver[i], h.lane_stream[i]) != CUSOLVER_STATUS_SUCCESS) {
        set_error("cusolverDnCreate (lane) failed");
        return TD_ERR_SOLVER;
      }
    }
    if (cublasCreate(&h.blas_side) != CUBLAS_STATUS_SUCCESS ||
        cublasSetStream(h.blas_side, h.lane_stream[0]) != CUBLAS_STATUS_SUCCESS) {
      set_error("cublasCreate (side) failed");
      return TD_ERR_SOLVER;
    }
  }
  *out = &h;
  return TD_OK;
}

}  // namespace

// td_shutdown: cuSOLVER / cuBLAS handles, side streams and events of every device that used a solve.
void release_solver_handles() {
  int cur = 0;
  cudaGetDevice(&cur);
  for (int dev = 0; dev < 64; ++dev) {
    Handles& h = g_handles[dev];
    if (!h.solver) continue;
    cudaSetDevice(dev);
    for (int i = 0; i < MAX_LANES; ++i) {
      if (h.lane_solver[i]) cusolverDnDestroy(h.lane_solver[i]);
      if (h.ev_done[i]) cudaEventDestroy(h.ev_done[i]);
      if (h.lane_stream[i]) cudaStreamDestroy(h.lane_stream[i]);
    }
    for (int i = 0; i < 8; ++i) {
      if (h.ev_panel[i]) cudaEventDestroy(h.ev_panel[i]);
      if (h.ev_wide[i]) cudaEventDestroy(h.ev_wide[i]);
    }
    if (h.ev_ready) cudaEventDestroy(h.ev_ready);
    if (h.blas_side) cublasDestroy(h.blas_side);
    if (h.blas) cublasDestroy(h.blas);
    cusolverDnDestroy(h.solver);
    h = Handles();
  }
  cudaSetDevice(cur);
}

namespace {

#define TD_CHECK_SOLVER(expr)                                                         \
  do {                                                                                \
    int s__ = (int)(expr);                                                            \
    if (s__ != 0) {                                                                   \
      td::set_error("%s failed with status %d (%s:%d)", #expr, s__, __FILE__, __LINE__); \
      return TD_ERR_SOLVER;                                                           \
    }                                                                                 \
  } while (0)

struct Scratch {   // stream-ordered device scratch, released on scope exit
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : st(s) {}
  ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  template <typename T>
  int get(size_t n, T** p) {
    void* q = nullptr;
    if (int rc = ensure_mempool_config()) return rc;
    cudaError_t e = cudaMallocAsync(&q, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e != cudaSuccess) {
      set_error("cudaMallocAsync of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
      return TD_ERR_CUDA;
    }
    ptrs.push_back(q);
    *p = static_cast<T*>(q);
    return TD_OK;
  }
};

// out = in * scale  (+ diag on the diagonal), n x n
__global__ void scale_add_diag(const double* in, double* out, long long n, double scale, double diag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n) return;
  const long long r = i / n, c = i - r * n;
  out[i] = in[i] * scale + (r == c ? diag : 0.0);
}
// row-major [rows, cols] -> column-major [rows, cols] with scale
__global__ void to_colmajor(const double* in, double* out, long long rows, long long cols, double scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols, c = i - r * cols;
  out[c * rows + r] = in[i] * scale;
}
// column-major double [rows, cols] (ld) -> row-major float [rows, cols]
__global__ void to_rowmajor_f32(const double* in, long long ld, float* out, long long rows, long long cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols, c = i - r * cols;
  out[i] = (float)in[c * ld + r];
}
__global__ void divide_rows(const double* z, const double* w, double lambda, double* out, long long rows, long long cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // column-major [rows, cols]
  if (i >= rows * cols) return;
  const long long r = i % rows;
  out[i] = z[i] / (w[r] + lambda);
}
// cov = G / norm - mu_a mu_b' (+ reg on the diagonal when square), row-major [ra, rb]
__global__ void make_cov(const double* G, const double* sa, const double* sb, double frames, double norm,
                         double reg, int add_reg, double* out, long long ra, long long rb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ra * rb) return;
  const long long r = i / rb, c = i - r * rb;
  double v = G[i] / norm - (sa[r] / frames) * (sb[c] / frames);
  if (add_reg && r == c) v += reg;
  out[i] = v;
}
// V[:, j] *= (w_j > eps ? w_j^-1/4 : 0), column-major n x n
__global__ void scale_cols_inv4(double* V, const double* w, double eps, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n) return;
  const long long j = i / n;
  const double wj = w[j];
  V[i] *= (wj > eps) ? 1.0 / sqrt(sqrt(wj)) : 0.0;
}
__global__ void mean_f32(const double* s, double frames, float* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(s[i] / frames);
}
__global__ void copy_f32(const double* in, float* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}
// out[:, i] = Q[:, n-1-i] for i < dim (eigenvectors of the `dim` largest eigenvalues, largest first),
// sigma[i] = sqrt(max(w[n-1-i], 0)); Q column-major [n, n], out column-major [n, dim].
__global__ void take_top_eigvecs(const double* Q, const double* w, long long n, int dim, double* out, double* sigma) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) sigma[i] = sqrt(fmax(w[n - 1 - i], 0.0));
  if (i >= n * dim) return;
  const long long c = i / n, r = i - c * n;
  out[i] = Q[(n - 1 - c) * n + r];
}
// V[:, j] /= sigma[j], column-major [rows, cols]
__global__ void scale_cols_recip(double* V, const double* sigma, long long rows, long long cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  V[i] /= sigma[i / rows];
}
// One block per column i of the column-major pair (rx [D1, dim], ry [D2, dim]): make the largest
// |entry| of rx[:, i] positive, flipping both columns together.
__global__ void normalise_signs(double* rx, long long D1, double* ry, long long D2) {
  __shared__ double best_abs[256];
  __shared__ double best_val[256];
  double* cx = rx + (long long)blockIdx.x * D1;
  double* cy = ry + (long long)blockIdx.x * D2;
  double ba = -1.0, bv = 0.0;
  for (long long r = threadIdx.x; r < D1; r += blockDim.x) {
    const double v = cx[r];
    if (fabs(v) > ba) { ba = fabs(v); bv = v; }
  }
  best_abs[threadIdx.x] = ba;
  best_val[threadIdx.x] = bv;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o && best_abs[threadIdx.x + o] > best_abs[threadIdx.x]) {
      best_abs[threadIdx.x] = best_abs[threadIdx.x + o];
      best_val[threadIdx.x] = best_val[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (best_val[0] < 0.0) {
    for (long long r = threadIdx.x; r < D1; r += blockDim.x) cx[r] = -cx[r];
    for (long long r = threadIdx.x; r < D2; r += blockDim.x) cy[r] = -cy[r];
  }
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

// Symmetric eigendecomposition in place (column-major == row-major for symmetric input).
int eigh(Handles* h, Scratch& sc, double* A, int n, double* w, int* info_dev) {
  int lwork = 0;
  TD_CHECK_SOLVER(cusolverDnDsyevd_bufferSize(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, &lwork));
  double* work;
  int rc = sc.get((size_t)lwork, &work);
  if (rc) return rc;
  TD_CHECK_SOLVER(cusolverDnDsyevd(h->solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, w, work, lwork, info_dev));
  return TD_OK;
}

// ---------------------------------------------------------------------------------------------
// Whole regularisation sweep as ONE batched, blocked Cholesky (float64).
//
// Every lambda gets its own copy M_l of the augmented matrix [[A0 + lambda I, .], [B0', .]]
// (column-major, Nt = Da + Dy rows; the right-hand sides ride along as extra rows, so the forward
// substitution L^-1 B falls out of the panel solves).  Per 64-column panel, for all lambdas at once:
//   potf2 of the diagonal block (one CTA per matrix, shared memory)
//   panel rows below:  P <- P L_jj^-T          (thread per row, L_jj in shared memory)
//   trailing update:   C <- C - P P'           (cublasDgemmStridedBatched over the batch)
// then a blocked back substitution.  One n ~ 2000 potrf cannot fill 148 SMs; twenty together can.
constexpr int CH_NB = 64;

__global__ void chol_build(const double* __restrict__ A0, const double* __restrict__ B0, int Da, int Dy,
                           const double* __restrict__ lambdas_dev, int l0, double* __restrict__ M) {
  const long long Nt = Da + Dy;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Nt * Da) return;                       // columns >= Da are never read
  const long long c = i / Nt, r = i - c * Nt;
  double* Ml = M + (size_t)blockIdx.y * Nt * Nt;
  double v;
  if (r < Da) v = A0[c * Da + r] + (r == c ? lambdas_dev[l0 + blockIdx.y] : 0.0);   // A0 is symmetric
  else v = B0[c + (r - Da) * Da];
  Ml[c * Nt + r] = v;
}

// The same working matrices for a batch of DIFFERENT systems (leave-one-trial-out folds): entry b is
// [[G_b / n_b + lambda_b I, .], [C_b' / n_b, .]] from row-major G [batch][Da][Da] (symmetric) and
// C [batch][Da][Dy].
__global__ void chol_build_folds(const double* __restrict__ G, const double* __restrict__ C, int Da, int Dy,
                                 const double* __restrict__ inv_n_dev, const double* __restrict__ lambdas_dev,
                                 int l0, double* __restrict__ M) {
  const long long Nt = Da + Dy;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Nt * Da) return;
  const long long c = i / Nt, r = i - c * Nt;
  const int b = l0 + blockIdx.y;
  double* Ml = M + (size_t)blockIdx.y * Nt * Nt;
  const double inv_n = inv_n_dev[b];
  double v;
  if (r < Da) v = G[(size_t)b * Da * Da + c * Da + r] * inv_n + (r == c ? lambdas_dev[b] : 0.0);
  else v = C[(size_t)b * Da * Dy + c * Dy + (r - Da)] * inv_n;
  Ml[c * Nt + r] = v;
}

// Factor the jb x jb diagonal block at (j0, j0) of every working matrix M, store the factor (lower)
// at the same place of L and write the inverse of the factor, Linv [batch][64][64] column-major (zero padded), for the panel solve.
// info[l] = first non-positive pivot (1-based, global column) or unchanged.
//
// This kernel sits on the serial chain of the blocked factorisation (once per 64 columns, one CTA per
// matrix), so it is built for latency, not throughput: the block is processed in four 16-column
// steps; the 16 x 16 diagonal piece is factored AND inverted by one warp entirely in registers
// (row per lane, columns and pivots exchanged by shuffles, no block-wide barrier inside), the panel
// below it and the trailing update are short dot products from shared memory by all 256 threads
// (three barriers per step).  The off-diagonal blocks of the inverse follow by block distance.
constexpr int CH_SB = 16;                        // sub-block of chol_potf2

__device__ __forceinline__ double shfl_f64(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

__global__ void __launch_bounds__(256) chol_potf2(const double* __restrict__ M, double* __restrict__ L, long long Nt,
                                                  int j0, int jb, int* info, double* __restrict__ Linv,
                                                  long long linv_stride) {
  extern __shared__ double chol_sm[];
  double (*T)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(chol_sm);                          // the block: lower triangle, identity padding
  double (*X)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(chol_sm + CH_NB * (CH_NB + 1));    // its inverse
  double (*P)[CH_SB + 1] = reinterpret_cast<double (*)[CH_SB + 1]>(chol_sm + 2 * CH_NB * (CH_NB + 1)); // panel of the current step
  const double* Ml = M + (size_t)blockIdx.x * Nt * Nt + (size_t)j0 * Nt + j0;
  double* Ll = L + (size_t)blockIdx.x * Nt * Nt + (size_t)j0 * Nt + j0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int c = e >> 6, r = e & 63;
    T[r][c] = (r >= c && r < jb) ? Ml[(size_t)c * Nt + r] : ((r == c) ? 1.0 : 0.0);
    X[r][c] = 0.0;
  }
  __syncthreads();
  for (int o = 0; o < CH_NB; o += CH_SB) {
    if (warp == 0) {
      // ---- 16 x 16 diagonal piece: lane i (< 16) owns row i; lanes >= 16 mirror lane i - 16 ----
      const int i = lane & 15;
      double a[CH_SB];
#pragma unroll
      for (int c = 0; c < CH_SB; ++c) a[c] = (c <= i) ? T[o + i][o + c] : 0.0;
      int bad = 0;
#pragma unroll
      for (int k = 0; k < CH_SB; ++k) {
        const double d = shfl_f64(a[k], k);
        if (!(d > 0.0) && bad == 0) bad = o + k + 1;
        // 1/sqrt(d): float seed + two Newton steps in double (relative error ~1e-16)
        const double dd = d > 0.0 ? d : 1.0;
        double rp = (double)rsqrtf((float)dd);
        rp = rp * (1.5 - 0.5 * dd * rp * rp);
        rp = rp * (1.5 - 0.5 * dd * rp * rp);
        const double lik = (i >= k) ? a[k] * rp : 0.0;       // L[i][k]; sqrt(d) on the diagonal
        a[k] = lik;
#pragma unroll
        for (int c = k + 1; c < CH_SB; ++c) {
          const double lck = shfl_f64(lik, c);
          if (c <= i) a[c] -= lik * lck;
        }
      }
      if (lane == 0 && bad && bad <= jb && info[blockIdx.x] == 0) info[blockIdx.x] = j0 + bad;
      // inverse of the piece: lane j builds column j of Xd = Ld^-1 by forward substitution;
      // L[r][k] is register a[k] of lane r, broadcast by shuffle (uniform r, k).
      double x[CH_SB];
      const double inv_diag = 1.0 / a[i];                    // 1 / L[i][i]
#pragma unroll
      for (int r = 0; r < CH_SB; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < r; ++k) acc += shfl_f64(a[k], r) * x[k];
        const double dr = shfl_f64(inv_diag, r);
        x[r] = (r < i) ? 0.0 : ((r == i) ? dr : -acc * dr);
      }
      if (lane < CH_SB) {
#pragma unroll
        for (int c = 0; c < CH_SB; ++c) {
          if (c <= i) T[o + i][o + c] = a[c];
          X[o + c][o + i] = x[c];                            // X[row c][col i] (zero above the diagonal)
        }
      }
    }
    __syncthreads();
    const int below = CH_NB - o - CH_SB;                     // rows under the piece
    // ---- panel: P[r][c] = sum_k T[o+16+r][o+k] * Xd[c][k]   (= A21 * Ld^-T) ----
    for (int e = tid; e < below * CH_SB; e += 256) {
      const int r = e >> 4, c = e & 15;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < CH_SB; ++k) acc += T[o + CH_SB + r][o + k] * X[o + c][o + k];
      P[r][c] = acc;
    }
    __syncthreads();
    for (int e = tid; e < below * CH_SB; e += 256) {
      const int r = e >> 4, c = e & 15;
      T[o + CH_SB + r][o + c] = P[r][c];
    }
    // ---- trailing update of the lower triangle: T[r][c] -= sum_k P[r][k] * P[c][k] ----
    for (int e = tid; e < below * below; e += 256) {
      const int r = e / below, c = e - r * below;
      if (c <= r) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < CH_SB; ++k) acc += P[r][k] * P[c][k];
        T[o + CH_SB + r][o + CH_SB + c] -= acc;
      }
    }
    __syncthreads();
  }
  // ---- off-diagonal blocks of X = L^-1, by block distance s:
  //      X[i][j] = -Xd_i * sum_{k=j}^{i-1} L[i][k] X[k][j],  i = j + s  (16 x 16 blocks) ----
  constexpr int NBK = CH_NB / CH_SB;
  for (int s = 1; s < NBK; ++s) {
    double keep[NBK - 1];
#pragma unroll
    for (int q = 0; q < NBK - 1; ++q) {                      // block column j = q (if j + s < 4)
      keep[q] = 0.0;
      if (q + s < NBK) {
        const int r = tid >> 4, c = tid & 15, bi = q + s;
        double acc = 0.0;
        for (int k = q * CH_SB; k < bi * CH_SB; ++k) acc += T[bi * CH_SB + r][k] * X[k][q * CH_SB + c];
        keep[q] = acc;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NBK - 1; ++q)
      if (q + s < NBK) X[(q + s) * CH_SB + (tid >> 4)][q * CH_SB + (tid & 15)] = keep[q];
    __syncthreads();
    // left-multiply by -Xd_i (in place needs a copy of the S block: read all, barrier, write)
#pragma unroll
    for (int q = 0; q < NBK - 1; ++q) {
      keep[q] = 0.0;
      if (q + s < NBK) {
        const int r = tid >> 4, c = tid & 15, bi = q + s;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < CH_SB; ++k) acc += X[bi * CH_SB + r][bi * CH_SB + k] * X[bi * CH_SB + k][q * CH_SB + c];
        keep[q] = -acc;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NBK - 1; ++q)
      if (q + s < NBK) X[(q + s) * CH_SB + (tid >> 4)][q * CH_SB + (tid & 15)] = keep[q];
    __syncthreads();
  }
  double* Li = Linv + (size_t)blockIdx.x * linv_stride + (size_t)(j0 / CH_NB) * CH_NB * CH_NB;
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int c = e >> 6, r = e & 63;
    Li[e] = X[r][c];
    if (r >= c && r < jb && c < jb) Ll[(size_t)c * Nt + r] = T[r][c];
  }
}

// Back substitution step for block j: x_j = L_jj^-T y_j (a 64 x 64 product with the stored inverse,
// recomputed by every CTA), then y[0:j0] -= L[j0:j0+jb, 0:j0]' x_j, one thread per column.
// Y: [batch][Dy][Da] work vectors.  grid (ceil(max(j0,1)/256), batch, Dy); CTA x == 0 stores x_j.
__global__ void __launch_bounds__(256) chol_back(const double* __restrict__ M, long long Nt, int Da, int j0,
                                                 int jb, const double* __restrict__ Linv, long long linv_stride,
                                                 double* Y, double* __restrict__ Xs) {
  __shared__ double yv[CH_NB];
  __shared__ double xv[CH_NB];
  const double* Ml = M + (size_t)blockIdx.y * Nt * Nt;
  const double* Li = Linv + (size_t)blockIdx.y * linv_stride + (size_t)(j0 / CH_NB) * CH_NB * CH_NB;
  double* y = Y + ((size_t)blockIdx.y * gridDim.z + blockIdx.z) * Da;
  if (threadIdx.x < CH_NB) yv[threadIdx.x] = (threadIdx.x < jb) ? y[j0 + threadIdx.x] : 0.0;
  __syncthreads();
  if (threadIdx.x < CH_NB) {             // x_i = sum_r Linv[r][i] y_r   (Linv column-major: Li[i*64 + r])
    const double* col = Li + (size_t)threadIdx.x * CH_NB;
    double sacc = 0.0;
#pragma unroll 8
    for (int r = 0; r < CH_NB; ++r) sacc += col[r] * yv[r];
    xv[threadIdx.x] = sacc;
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < j0) {
    const double* col = Ml + (size_t)c * Nt + j0;
    double sacc = 0.0;
#pragma unroll 8
    for (int r = 0; r < jb; ++r) sacc += col[r] * xv[r];
    y[c] -= sacc;
  }
  // The solution goes to a separate vector: other CTAs of this launch still read y_j.
  if (blockIdx.x == 0 && threadIdx.x < jb)
    Xs[((size_t)blockIdx.y * gridDim.z + blockIdx.z) * Da + j0 + threadIdx.x] = xv[threadIdx.x];
}

__global__ void chol_take_y(const double* __restrict__ M, long long Nt, int Da, int Dy, double* Y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Da * Dy) return;
  const int q = (int)(i / Da), c = (int)(i - (long long)q * Da);
  Y[((size_t)blockIdx.y * Dy + q) * Da + c] = M[(size_t)blockIdx.y * Nt * Nt + (size_t)c * Nt + Da + q];
}

__global__ void chol_store(const double* __restrict__ Y, int Da, int Dy, float* W) {   // W [batch][Da][Dy]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Da * Dy) return;
  const int r = (int)(i / Dy), q = (int)(i - (long long)r * Dy);
  W[(size_t)blockIdx.y * Da * Dy + i] = (float)Y[((size_t)blockIdx.y * Dy + q) * Da + r];
}

// folds_inv_n != nullptr: A0 / B0 are the batched fold inputs G [n_lambda][Da][Da], C [n_lambda][Da][Dy]
// (row-major, un-normalised) and folds_inv_n [n_lambda] (host) their 1 / n.
int batched_cholesky_sweep(Handles* h, Scratch& sc, cudaStream_t st, const double* A0, const double* B0,
                           int Da, int Dy, const double* lambdas, int n_lambda, float* W_out,
                           std::vector<int>* failed, const double* folds_inv_n = nullptr) {
  int rc;
  const long long Nt = (long long)Da + Dy;
  const size_t per = (size_t)Nt * Nt * sizeof(double);
  const int batch_max = (int)std::max<size_t>(1, std::min<size_t>(n_lambda, ((size_t)6 << 30) / (2 * per)));
  const long long n_panels = (Da + CH_NB - 1) / CH_NB;
  const long long linv_stride = n_panels * CH_NB * CH_NB;
  double *M, *L, *Y, *Xs, *lam_dev, *Linv;     // M: working matrices (updated in place), L: the factors
  int* info;
  if ((rc = sc.get((size_t)Nt * Nt * batch_max, &M)) || (rc = sc.get((size_t)Da * Dy * batch_max, &Y)) || (rc = sc.get((size_t)Da * Dy * batch_max, &Xs)) ||
      (rc = sc.get(n_lambda, &lam_dev)) || (rc = sc.get(n_lambda, &info)) ||
      (rc = sc.get((size_t)linv_stride * batch_max, &Linv)) || (rc = sc.get((size_t)Nt * Nt * batch_max, &L)))
    return rc;
  TD_CHECK_CUDA(cudaMemcpyAsync(lam_dev, lambdas, n_lambda * sizeof(double), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemsetAsync(info, 0, n_lambda * sizeof(int), st));
  double* inv_n_dev = nullptr;
  if (folds_inv_n) {
    if ((rc = sc.get(n_lambda, &inv_n_dev))) return rc;
    TD_CHECK_CUDA(cudaMemcpyAsync(inv_n_dev, folds_inv_n, n_lambda * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  const double minus_one = -1.0, one = 1.0, zero = 0.0;
  const int potf2_smem = (2 * CH_NB * (CH_NB + 1) + CH_NB * (CH_SB + 1)) * (int)sizeof(double);
  TD_CHECK_CUDA(cudaFuncSetAttribute(chol_potf2, cudaFuncAttributeMaxDynamicSharedMemorySize, potf2_smem));
  const char* dbg_env = getenv("TD_DEBUG");
  const bool dbg = dbg_env && dbg_env[0] == '2';
  cudaEvent_t ev[4];
  if (dbg) for (auto& e : ev) cudaEventCreate(&e);
  for (int l0 = 0; l0 < n_lambda; l0 += batch_max) {
    const int nbat = std::min(batch_max, n_lambda - l0);
    if (dbg) cudaEventRecord(ev[0], st);
    if (folds_inv_n)
      chol_build_folds<<<dim3(blocks_for(Nt * Da), nbat), 256, 0, st>>>(A0, B0, Da, Dy, inv_n_dev, lam_dev, l0, M);
    else
      chol_build<<<dim3(blocks_for(Nt * Da), nbat), 256, 0, st>>>(A0, B0, Da, Dy, lam_dev, l0, M);
    if (dbg) cudaEventRecord(ev[1], st);
    // Four 64-column panels form a 256-column block (measured 2 / 4 / 8 panels: 4.67 / 4.43 / 4.74 ms for the
    // factorisation of 20 systems of n = 2049).  The serial chain -- per panel
    // potf2, panel solve and the rank-64 update of the block's remaining columns, then the rank-256
    // update of the NEXT block's columns -- runs on the caller's stream; the wide rank-256 trailing
    // update of block b (everything right of the next block) runs on a side stream and overlaps the
    // chain of block b + 1 (look-ahead).
    cudaStream_t side = h->lane_stream[0];
    auto panel_factor = [&](int j0, int jb) -> int {
      chol_potf2<<<nbat, 256, potf2_smem, st>>>(M, L, Nt, j0, jb, info + l0, Linv, linv_stride);
      const long long below = Nt - j0 - jb;
      if (below <= 0) return TD_OK;
      // panel solve L21 = A21 L_jj^-T as a GEMM with the explicit 64 x 64 inverse, out of place
      // (working matrix -> factor matrix), so no copy of the panel is needed
      TD_CHECK_SOLVER(cublasDgemmStridedBatched(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)below, jb, jb, &one,
                                                M + (size_t)j0 * Nt + j0 + jb, (int)Nt, Nt * Nt,
                                                Linv + (size_t)(j0 / CH_NB) * CH_NB * CH_NB, CH_NB,
                                                linv_stride, &zero, L + (size_t)j0 * Nt + j0 + jb,
                                                (int)Nt, Nt * Nt, nbat));
      return TD_OK;
    };
    // C[c0:, c0:c0+nc] -= P[c0:, :] P[c0:c0+nc, :]'  with P = L[:, p0:p0+k], C in the working matrix
    auto update = [&](cublasHandle_t blas, long long p0, int k, long long c0, long long nc) -> int {
      const double* Prow = L + (size_t)p0 * Nt + c0;
      double* C = M + (size_t)c0 * Nt + c0;
      TD_CHECK_SOLVER(cublasDgemmStridedBatched(blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)(Nt - c0), (int)nc, k,
                                                &minus_one, Prow, (int)Nt, Nt * Nt, Prow, (int)Nt, Nt * Nt, &one,
                                                C, (int)Nt, Nt * Nt, nbat));
      return TD_OK;
    };
#ifndef TD_CHOL_PANELS_PER_BLOCK
#define TD_CHOL_PANELS_PER_BLOCK 4
#endif
    constexpr int CH_BLK = TD_CHOL_PANELS_PER_BLOCK * CH_NB;
    int blk = 0;
    for (int b0 = 0; b0 < Da; b0 += CH_BLK, ++blk) {
      const int kb = std::min(CH_BLK, Da - b0);              // columns of this block
      // The block's columns got their last updates from the previous block's narrow update (this
      // stream) and from wide updates that the narrow update had already waited for.  Inside the
      // block: factor a 64-column panel, rank-64 update of the block's remaining columns, repeat.
      for (int j0 = b0; j0 < b0 + kb; j0 += CH_NB) {
        const int jb = std::min(CH_NB, b0 + kb - j0);
        if ((rc = panel_factor(j0, jb))) return rc;
        const int rest = b0 + kb - (j0 + jb);
        if (rest > 0 && (rc = update(h->blas, j0, jb, j0 + jb, rest))) return rc;
      }
      TD_CHECK_CUDA(cudaEventRecord(h->ev_panel[blk & 7], st));
      const long long t2 = (long long)b0 + kb;
      if (t2 >= Da) continue;
      // narrow update: the next block's columns, after the previous wide update (which also wrote
      // them) has finished
      if (blk > 0) TD_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_wide[(blk - 1) & 7], 0));
      const long long nnar = std::min<long long>(CH_BLK, Da - t2);
      if ((rc = update(h->blas, b0, kb, t2, nnar))) return rc;
      // wide update on the side stream: columns [t2 + 128, Da), lower block-triangle, 512-column blocks
      TD_CHECK_CUDA(cudaStreamWaitEvent(side, h->ev_panel[blk & 7], 0));
      for (long long cb = t2 + nnar; cb < Da; cb += 512)
        if ((rc = update(h->blas_side, b0, kb, cb, std::min<long long>(512, Da - cb)))) return rc;
      TD_CHECK_CUDA(cudaEventRecord(h->ev_wide[blk & 7], side));
    }
    const int panel = blk;
    // everything the side stream did must be visible before the back substitution
    if (panel > 0) TD_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_wide[(panel - 1) & 7], 0));
    TD_CHECK_CUDA(cudaEventRecord(h->ev_ready, side));
    TD_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_ready, 0));
    if (dbg) cudaEventRecord(ev[2], st);
    chol_take_y<<<dim3(blocks_for((long long)Da * Dy), nbat), 256, 0, st>>>(L, Nt, Da, Dy, Y);
    for (int j0 = ((Da - 1) / CH_NB) * CH_NB; j0 >= 0; j0 -= CH_NB) {
      const int jb = std::min(CH_NB, Da - j0);
      chol_back<<<dim3(blocks_for(std::max(j0, 1)), nbat, Dy), 256, 0, st>>>(L, Nt, Da, j0, jb, Linv, linv_stride, Y, Xs);
    }
    chol_store<<<dim3(blocks_for((long long)Da * Dy), nbat), 256, 0, st>>>(Xs, Da, Dy, W_out + (size_t)l0 * Da * Dy);
    TD_CHECK_CUDA(cudaGetLastError());
    if (dbg) {
      cudaEventRecord(ev[3], st);
      cudaEventSynchronize(ev[3]);
      float a, b, c;
      cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
      fprintf(stderr, "[tdgram] batched cholesky: build %.3f ms, factor %.3f ms, back-substitution %.3f ms (batch %d, n %d)\n", a, b, c, nbat, Da);
    }
  }
  std::vector<int> hinfo(n_lambda, 0);
  TD_CHECK_CUDA(cudaMemcpyAsync(hinfo.data(), info, n_lambda * sizeof(int), cudaMemcpyDeviceToHost, st));
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int l = 0; l < n_lambda; ++l)
    if (hinfo[l] != 0) failed->push_back(l);
  return TD_OK;
}

// Cholesky per lambda on already normalised A0 [Da, Da], B0 (column-major [Da, Dy]).
int td_ridge_solve_scaled(Handles* h, Scratch& sc, cudaStream_t st, const double* A0, const double* B0,
                          int Da, int Dy, const double* lambdas, int n_lambda, float* W_out) {
  int rc;
  const long long nn = (long long)Da * Da, nb = (long long)Da * Dy;
  double* A;
  if ((rc = sc.get(nn, &A))) return rc;
  // Cholesky per lambda.  The factorisations are independent: they run on up to MAX_LANES side
  // streams, nothing synchronises with the host until all are queued, and only a lambda whose
  // matrix turned out not to be positive definite is redone with LU (like np.linalg.solve).
  const int lanes = (Da > 8192) ? 1 : std::min(MAX_LANES, n_lambda);
  int lwork_p = 0, lwork_g = 0;
  TD_CHECK_SOLVER(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, Da, A, Da, &lwork_p));
  TD_CHECK_SOLVER(cusolverDnDgetrf_bufferSize(h->solver, Da, Da, A, Da, &lwork_g));
  double *lane_A[MAX_LANES], *lane_work[MAX_LANES], *Xall;
  int* infos;
  lane_A[0] = A;
  for (int i = 1; i < lanes; ++i)
    if ((rc = sc.get(nn, &lane_A[i]))) return rc;
  for (int i = 0; i < lanes; ++i)
    if ((rc = sc.get((size_t)std::max(lwork_p, lwork_g), &lane_work[i]))) return rc;
  if ((rc = sc.get((size_t)nb * n_lambda, &Xall)) || (rc = sc.get(2 * n_lambda, &infos))) return rc;
  TD_CHECK_CUDA(cudaEventRecord(h->ev_ready, st));
  for (int i = 0; i < lanes; ++i) TD_CHECK_CUDA(cudaStreamWaitEvent(h->lane_stream[i], h->ev_ready, 0));
  for (int l = 0; l < n_lambda; ++l) {
    const int i = l % lanes;
    cudaStream_t ls = h->lane_stream[i];
    double* Xl = Xall + (size_t)l * nb;
    scale_add_diag<<<blocks_for(nn), 256, 0, ls>>>(A0, lane_A[i], Da, 1.0, lambdas[l]);
    TD_CHECK_CUDA(cudaMemcpyAsync(Xl, B0, nb * sizeof(double), cudaMemcpyDeviceToDevice, ls));
    TD_CHECK_SOLVER(cusolverDnDpotrf(h->lane_solver[i], CUBLAS_FILL_MODE_LOWER, Da, lane_A[i], Da, lane_work[i], lwork_p, infos + 2 * l));
    TD_CHECK_SOLVER(cusolverDnDpotrs(h->lane_solver[i], CUBLAS_FILL_MODE_LOWER, Da, Dy, lane_A[i], Da, Xl, Da, infos + 2 * l + 1));
    to_rowmajor_f32<<<blocks_for(nb), 256, 0, ls>>>(Xl, Da, W_out + (size_t)l * nb, Da, Dy);
  }
  TD_CHECK_CUDA(cudaGetLastError());
  for (int i = 0; i < lanes; ++i) {
    TD_CHECK_CUDA(cudaEventRecord(h->ev_done[i], h->lane_stream[i]));
    TD_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_done[i], 0));
  }
  std::vector<int> hinfos(2 * n_lambda, 0);
  TD_CHECK_CUDA(cudaMemcpyAsync(hinfos.data(), infos, hinfos.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  int* ipiv = nullptr;
  for (int l = 0; l < n_lambda; ++l) {
    if (hinfos[2 * l] == 0) continue;
    // Not positive definite (lambda = 0 on a rank-deficient design): LU like np.linalg.solve.
    if (!ipiv && (rc = sc.get(Da, &ipiv))) return rc;
    double* Xl = Xall + (size_t)l * nb;
    scale_add_diag<<<blocks_for(nn), 256, 0, st>>>(A0, A, Da, 1.0, lambdas[l]);
    TD_CHECK_CUDA(cudaMemcpyAsync(Xl, B0, nb * sizeof(double), cudaMemcpyDeviceToDevice, st));
    TD_CHECK_SOLVER(cusolverDnDgetrf(h->solver, Da, Da, A, Da, lane_work[0], ipiv, infos));
    int hinfo = 0;
    TD_CHECK_CUDA(cudaMemcpyAsync(&hinfo, infos, sizeof(int), cudaMemcpyDeviceToHost, st));
    TD_CHECK_CUDA(cudaStreamSynchronize(st));
    if (hinfo != 0) {
      set_error("td_ridge_solve: singular matrix for lambda=%g (LU info %d)", lambdas[l], hinfo);   // numpy: LinAlgError("Singular matrix")
      return TD_ERR_SOLVER;
    }
    TD_CHECK_SOLVER(cusolverDnDgetrs(h->solver, CUBLAS_OP_N, Da, Dy, A, Da, ipiv, Xl, Da, infos));
    to_rowmajor_f32<<<blocks_for(nb), 256, 0, st>>>(Xl, Da, W_out + (size_t)l * nb, Da, Dy);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  return TD_OK;
}

// ---------------------------------------------------------------------------------------------
// Leave-one-trial-out with ONE factorisation.
//
// Fold f solves (G_tot - G_f + lambda n_f I) w = C_tot - C_f  (brain_model.py:447-454 on the fold's
// sums, n_f = n_tot - rows of trial f).  With S = G_tot + lambda n_ref I (n_ref = mean n_f) this is
// (S - E_f) w = c_f, E_f = G_f + lambda (n_ref - n_f) I -- a small perturbation of ONE matrix shared
// by every fold (a trial holds ~1/F of the energy), so after one Cholesky of S the fixed point
//     w <- S^-1 (c_f + E_f w)
// contracts by ~1/F per step: 3 - 5 steps to float64 accuracy, each one pass over the per-trial Grams
// (a memory-bound matrix-vector kernel) and one triangular solve with F*Dy right-hand sides, instead
// of F factorisations.  The caller falls back to the batched Cholesky when a fold does not converge
// (one trial dominating the recording).
// ---------------------------------------------------------------------------------------------
// R[:, f*Dy + q] = C_tot[:, q] - C_f[:, q]   (column-major Da x F*Dy);  C_f = [xtv_f[:, :Dy]; vtv_f[Dy, :Dy]]
__global__ void loo_build_rhs(const double* __restrict__ C_tot, const double* __restrict__ xtv, const double* __restrict__ vtv,
                              int D, int Dy, int F, double* __restrict__ R) {
  const int Da = D + 1, Q = Dy + 1;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Da * F * Dy) return;
  const long long col = i / Da;
  const int r = (int)(i - col * Da);
  const int f = (int)(col / Dy), q = (int)(col - (long long)f * Dy);
  const double cf = (r < D) ? xtv[((size_t)f * D + r) * Q + q] : vtv[(size_t)f * Q * Q + (size_t)Dy * Q + q];
  R[i] = C_tot[(size_t)r * Dy + q] - cf;
}

// T[:, f*Dy + q] = R[:, f*Dy + q] + E_f W[:, f*Dy + q]: one warp per (row, fold); gram [F][D][D] float32
// (symmetric), the ones column / row of the augmented Gram from xtv[:, :, Dy] and vtv[:, Dy, Dy].
template <int DY>
__global__ void __launch_bounds__(256) loo_apply(const float* __restrict__ gram, const double* __restrict__ xtv,
                                                 const double* __restrict__ vtv, const double* __restrict__ shift,
                                                 const double* __restrict__ W, const double* __restrict__ R,
                                                 int D, int F, double* __restrict__ T) {
  const int Da = D + 1, Q = DY + 1;
  const int f = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Da) return;
  const double* w = W + (size_t)f * DY * Da;
  double acc[DY];
#pragma unroll
  for (int q = 0; q < DY; ++q) acc[q] = 0.0;
  if (row < D) {
    const float* g = gram + ((size_t)f * D + row) * D;
    for (int j = lane; j < D; j += 32) {
      const double gv = (double)__ldg(g + j);
#pragma unroll
      for (int q = 0; q < DY; ++q) acc[q] += gv * w[(size_t)q * Da + j];
    }
  } else {
    const double* sf = xtv + (size_t)f * D * Q + DY;               // ones column: xtv[f][j][Dy]
    for (int j = lane; j < D; j += 32) {
      const double gv = sf[(size_t)j * Q];
#pragma unroll
      for (int q = 0; q < DY; ++q) acc[q] += gv * w[(size_t)q * Da + j];
    }
  }
#pragma unroll
  for (int q = 0; q < DY; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  if (lane == 0) {
    const double last = (row < D) ? xtv[((size_t)f * D + row) * Q + DY] : vtv[(size_t)f * Q * Q + (size_t)DY * Q + DY];
#pragma unroll
    for (int q = 0; q < DY; ++q) {
      const size_t o = ((size_t)f * DY + q) * Da + row;
      T[o] = R[o] + acc[q] + last * w[(size_t)q * Da + D] + shift[f] * w[(size_t)q * Da + row];
    }
  }
}

// per fold: max |Wn - W| and max |Wn| over its Da x Dy block (column-major, contiguous)
__global__ void loo_diff(const double* __restrict__ Wn, const double* __restrict__ W, long long per_fold, double* out) {
  __shared__ double sd[256], sm[256];
  const double* a = Wn + (size_t)blockIdx.x * per_fold;
  const double* b = W + (size_t)blockIdx.x * per_fold;
  double d = 0.0, m = 0.0;
  for (long long i = threadIdx.x; i < per_fold; i += blockDim.x) {
    d = fmax(d, fabs(a[i] - b[i]));
    m = fmax(m, fabs(a[i]));
  }
  sd[threadIdx.x] = d; sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sd[threadIdx.x] = fmax(sd[threadIdx.x], sd[threadIdx.x + o]); sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = sd[0]; out[2 * blockIdx.x + 1] = sm[0]; }
}

// column-major [Da, F*Dy] double -> W_out [F][Da][Dy] float
__global__ void loo_store(const double* __restrict__ W, int Da, int Dy, int F, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)F * Da * Dy) return;
  const int q = (int)(i % Dy);
  const long long t = i / Dy;
  const int r = (int)(t % Da);
  const int f = (int)(t / Da);
  out[i] = (float)W[((size_t)f * Dy + q) * Da + r];
}

}  // namespace
}  // namespace td

using namespace td;

extern "C" int td_ridge_solve(const double* G_aug, const double* C_aug, int Da, int Dy, double n,
                              const double* lambdas, int n_lambda, int method, float* W_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(G_aug && C_aug && lambdas && W_out, "td_ridge_solve: null argument");
  TD_REQUIRE(Da >= 1 && Dy >= 1 && n_lambda >= 1 && n > 0, "td_ridge_solve: Da=%d Dy=%d n_lambda=%d n=%g", Da, Dy, n_lambda, n);
  TD_REQUIRE(method >= 0 && method <= 2, "td_ridge_solve: unknown method %d", method);
  Handles* h;
  int rc = get_handles(&h);
  if (rc) return rc;
  TD_CHECK_SOLVER(cusolverDnSetStream(h->solver, st));
  TD_CHECK_SOLVER(cublasSetStream(h->blas, st));
  Scratch sc(st);
  const long long nn = (long long)Da * Da, nb = (long long)Da * Dy;
  double *A0, *A, *B0, *X;
  int* info;
  if ((rc = sc.get(nn, &A0)) || (rc = sc.get(nn, &A)) || (rc = sc.get(nb, &B0)) || (rc = sc.get(nb, &X)) || (rc = sc.get(1, &info))) return rc;
  scale_add_diag<<<blocks_for(nn), 256, 0, st>>>(G_aug, A0, Da, 1.0 / n, 0.0);
  to_colmajor<<<blocks_for(nb), 256, 0, st>>>(C_aug, B0, Da, Dy, 1.0 / n);
  TD_CHECK_CUDA(cudaGetLastError());

  if (method == 1) {
    double *w, *Z;
    if ((rc = sc.get(Da, &w)) || (rc = sc.get(nb, &Z))) return rc;
    if ((rc = eigh(h, sc, A0, Da, w, info))) return rc;    // A0 now holds V (column-major)
    int hinfo = 0;
    TD_CHECK_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
    TD_CHECK_CUDA(cudaStreamSynchronize(st));
    if (hinfo != 0) { set_error("td_ridge_solve: syevd did not converge (info %d)", hinfo); return TD_ERR_SOLVER; }
    const double one = 1.0, zero = 0.0;
    TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, Da, Dy, Da, &one, A0, Da, B0, Da, &zero, Z, Da));
    for (int l = 0; l < n_lambda; ++l) {
      divide_rows<<<blocks_for(nb), 256, 0, st>>>(Z, w, lambdas[l], A /*scratch*/, Da, Dy);
      TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, Da, Dy, Da, &one, A0, Da, A, Da, &zero, X, Da));
      to_rowmajor_f32<<<blocks_for(nb), 256, 0, st>>>(X, Da, W_out + (size_t)l * nb, Da, Dy);
    }
    TD_CHECK_CUDA(cudaGetLastError());
    TD_CHECK_CUDA(cudaStreamSynchronize(st));
    return TD_OK;
  }

  if (method == 2) {
    std::vector<int> failed;
    rc = batched_cholesky_sweep(h, sc, st, A0, B0, Da, Dy, lambdas, n_lambda, W_out, &failed);
    if (rc) return rc;
    if (failed.empty()) return TD_OK;
    // Not positive definite for some lambda: those go through the per-lambda path below (LU).
    std::vector<double> lam2;
    for (int l : failed) lam2.push_back(lambdas[l]);
    float* W2;
    if ((rc = sc.get((size_t)nb * failed.size(), &W2))) return rc;
    // (A0 / B0 are already normalised: pass n = 1 through a recursive call on the scaled data.)
    rc = td_ridge_solve_scaled(h, sc, st, A0, B0, Da, Dy, lam2.data(), (int)lam2.size(), W2);
    if (rc) return rc;
    for (size_t i = 0; i < failed.size(); ++i)
      TD_CHECK_CUDA(cudaMemcpyAsync(W_out + (size_t)failed[i] * nb, W2 + i * nb, nb * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TD_CHECK_CUDA(cudaStreamSynchronize(st));
    return TD_OK;
  }
  rc = td_ridge_solve_scaled(h, sc, st, A0, B0, Da, Dy, lambdas, n_lambda, W_out);
  if (rc) return rc;
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  return TD_OK;
}

// Leave-one-trial-out through one shared factorisation (see loo_* above).  G_aug_total [Da][Da] and
// C_aug_total [Da][Dy] double (Da = D + 1, ones column last), per-fold statistics of the held-out
// trials: gram_fold [F][D][D] float32, xtv_fold [F][D][Dy+1], vtv_fold [F][Dy+1][Dy+1] double,
// n_fold (HOST, rows of each held-out trial).  W_out [F][Da][Dy] float32.  worst_rel (HOST) receives
// the largest relative change of the last iteration over the folds; status TD_OK even if it is above
// tol -- the caller decides (fallback to td_ridge_solve_batched).  Synchronises the stream.
extern "C" int td_ridge_loo_solve(const double* G_aug_total, const double* C_aug_total, const float* gram_fold,
                                  const double* xtv_fold, const double* vtv_fold, const double* n_fold, int D,
                                  int Dy, int F, double n_total, double lambda, int max_iter, double tol,
                                  float* W_out, int* iterations, double* worst_rel, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(G_aug_total && C_aug_total && gram_fold && xtv_fold && vtv_fold && n_fold && W_out,
             "td_ridge_loo_solve: null argument");
  TD_REQUIRE(D >= 1 && F >= 1 && Dy >= 1 && Dy <= 4, "td_ridge_loo_solve: D=%d F=%d Dy=%d (Dy must be 1..4)", D, F, Dy);
  TD_REQUIRE(n_total > 0 && max_iter >= 1, "td_ridge_loo_solve: n_total=%g max_iter=%d", n_total, max_iter);
  Handles* h;
  int rc = get_handles(&h);
  if (rc) return rc;
  TD_CHECK_SOLVER(cusolverDnSetStream(h->solver, st));
  Scratch sc(st);
  const int Da = D + 1;
  const long long nn = (long long)Da * Da, ncol = (long long)F * Dy, nb = (long long)Da * ncol;
  double n_ref = 0.0;
  std::vector<double> shift(F);
  for (int f = 0; f < F; ++f) {
    TD_REQUIRE(n_fold[f] > 0 && n_fold[f] < n_total, "td_ridge_loo_solve: trial %d has %g of %g rows", f, n_fold[f], n_total);
    n_ref += (n_total - n_fold[f]) / F;
  }
  for (int f = 0; f < F; ++f) shift[f] = lambda * (n_ref - (n_total - n_fold[f]));     // E_f = G_f + shift_f I
  double *S, *R, *W, *T, *work, *shift_dev, *diff_dev;
  int* info;
  int lwork = 0;
  if ((rc = sc.get(nn, &S)) || (rc = sc.get(nb, &R)) || (rc = sc.get(nb, &W)) || (rc = sc.get(nb, &T)) ||
      (rc = sc.get(F, &shift_dev)) || (rc = sc.get(2 * F, &diff_dev)) || (rc = sc.get(2, &info)))
    return rc;
  TD_CHECK_CUDA(cudaMemcpyAsync(shift_dev, shift.data(), F * sizeof(double), cudaMemcpyHostToDevice, st));
  scale_add_diag<<<blocks_for(nn), 256, 0, st>>>(G_aug_total, S, Da, 1.0, lambda * n_ref);
  TD_CHECK_SOLVER(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, Da, S, Da, &lwork));
  if ((rc = sc.get((size_t)lwork, &work))) return rc;
  TD_CHECK_SOLVER(cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, Da, S, Da, work, lwork, info));
  loo_build_rhs<<<blocks_for(nb), 256, 0, st>>>(C_aug_total, xtv_fold, vtv_fold, D, Dy, F, R);
  TD_CHECK_CUDA(cudaMemcpyAsync(W, R, nb * sizeof(double), cudaMemcpyDeviceToDevice, st));
  TD_CHECK_SOLVER(cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, Da, (int)ncol, S, Da, W, Da, info + 1));
  int hinfo[2] = {0, 0};
  TD_CHECK_CUDA(cudaMemcpyAsync(hinfo, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  if (hinfo[0] != 0 || hinfo[1] != 0) {
    set_error("td_ridge_loo_solve: the total system is not positive definite (potrf info %d)", hinfo[0]);
    return TD_ERR_SOLVER;
  }
  std::vector<double> hd(2 * F);
  double worst = 0.0;
  int it = 0;
  for (it = 1; it <= max_iter; ++it) {
    dim3 grid((unsigned)((Da + 7) / 8), (unsigned)F);
    switch (Dy) {
      case 1: loo_apply<1><<<grid, 256, 0, st>>>(gram_fold, xtv_fold, vtv_fold, shift_dev, W, R, D, F, T); break;
      case 2: loo_apply<2><<<grid, 256, 0, st>>>(gram_fold, xtv_fold, vtv_fold, shift_dev, W, R, D, F, T); break;
      case 3: loo_apply<3><<<grid, 256, 0, st>>>(gram_fold, xtv_fold, vtv_fold, shift_dev, W, R, D, F, T); break;
      default: loo_apply<4><<<grid, 256, 0, st>>>(gram_fold, xtv_fold, vtv_fold, shift_dev, W, R, D, F, T); break;
    }
    TD_CHECK_CUDA(cudaGetLastError());
    TD_CHECK_SOLVER(cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, Da, (int)ncol, S, Da, T, Da, info + 1));
    loo_diff<<<F, 256, 0, st>>>(T, W, (long long)Da * Dy, diff_dev);
    TD_CHECK_CUDA(cudaMemcpyAsync(hd.data(), diff_dev, 2 * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    std::swap(W, T);
    TD_CHECK_CUDA(cudaStreamSynchronize(st));
    worst = 0.0;
    for (int f = 0; f < F; ++f) worst = std::max(worst, hd[2 * f] / std::max(hd[2 * f + 1], 1e-300));
    if (worst < tol) break;
  }
  loo_store<<<blocks_for((long long)F * Da * Dy), 256, 0, st>>>(W, Da, Dy, F, W_out);
  TD_CHECK_CUDA(cudaGetLastError());
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  if (iterations) *iterations = std::min(it, max_iter);
  if (worst_rel) *worst_rel = worst;
  return TD_OK;
}

// Leave-one-trial-out: a batch of different systems, one lambda each, in one blocked Cholesky.
extern "C" int td_ridge_solve_batched(const double* G_aug, const double* C_aug, int Da, int Dy,
                                      const double* n, const double* lambdas, int n_batch, float* W_out,
                                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(G_aug && C_aug && n && lambdas && W_out, "td_ridge_solve_batched: null argument");
  TD_REQUIRE(Da >= 1 && Dy >= 1 && n_batch >= 1, "td_ridge_solve_batched: Da=%d Dy=%d n_batch=%d", Da, Dy, n_batch);
  Handles* h;
  int rc = get_handles(&h);
  if (rc) return rc;
  TD_CHECK_SOLVER(cusolverDnSetStream(h->solver, st));
  TD_CHECK_SOLVER(cublasSetStream(h->blas, st));
  Scratch sc(st);
  std::vector<double> inv_n(n_batch);
  for (int b = 0; b < n_batch; ++b) {
    TD_REQUIRE(n[b] > 0, "td_ridge_solve_batched: system %d has n = %g rows", b, n[b]);
    inv_n[b] = 1.0 / n[b];
  }
  std::vector<int> failed;
  rc = batched_cholesky_sweep(h, sc, st, G_aug, C_aug, Da, Dy, lambdas, n_batch, W_out, &failed, inv_n.data());
  if (rc) return rc;
  if (!failed.empty()) {
    set_error("td_ridge_solve_batched: system %d of %d is not positive definite (lambda = %g)", failed[0], n_batch,
              lambdas[failed[0]]);
    return TD_ERR_SOLVER;
  }
  return TD_OK;
}

extern "C" int td_cca_solve(const double* Gxx, const double* Gyy, const double* Gxy, const double* sum_x,
                            const double* sum_y, int D1, int D2, double n_frames, double n_norm, double reg,
                            double eps_eig, int dim, float* rot_x, float* rot_y, float* mean_x, float* mean_y,
                            float* e_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(Gxx && Gyy && Gxy && sum_x && sum_y && rot_x && rot_y && mean_x && mean_y && e_out, "td_cca_solve: null argument");
  TD_REQUIRE(D1 >= 1 && D2 >= 1, "td_cca_solve: D1=%d D2=%d", D1, D2);
  TD_REQUIRE(dim >= 1 && dim <= std::min(D1, D2), "td_cca_solve: dim=%d must be in [1, min(D1, D2)=%d]", dim, std::min(D1, D2));
  TD_REQUIRE(reg >= 0.0, "regularization lambda must be >= 0");                 // cca.py:298-299
  TD_REQUIRE(n_frames > 0 && n_norm > 0, "td_cca_solve: n_frames=%g n_norm=%g", n_frames, n_norm);
  Handles* h;
  int rc = get_handles(&h);
  if (rc) return rc;
  TD_CHECK_SOLVER(cusolverDnSetStream(h->solver, st));
  TD_CHECK_SOLVER(cublasSetStream(h->blas, st));
  Scratch sc(st);
  const long long n11 = (long long)D1 * D1, n22 = (long long)D2 * D2, n12 = (long long)D1 * D2;
  double *Sxx, *Syy, *Sxy, *w1, *w2, *K11, *K22, *tmp, *Tt;
  int* info;
  if ((rc = sc.get(n11, &Sxx)) || (rc = sc.get(n22, &Syy)) || (rc = sc.get(n12, &Sxy)) || (rc = sc.get(D1, &w1)) ||
      (rc = sc.get(D2, &w2)) || (rc = sc.get(n11, &K11)) || (rc = sc.get(n22, &K22)) || (rc = sc.get(n12, &tmp)) ||
      (rc = sc.get(n12, &Tt)) || (rc = sc.get(2, &info)))
    return rc;
  make_cov<<<blocks_for(n11), 256, 0, st>>>(Gxx, sum_x, sum_x, n_frames, n_norm, reg, 1, Sxx, D1, D1);
  make_cov<<<blocks_for(n22), 256, 0, st>>>(Gyy, sum_y, sum_y, n_frames, n_norm, reg, 1, Syy, D2, D2);
  make_cov<<<blocks_for(n12), 256, 0, st>>>(Gxy, sum_x, sum_y, n_frames, n_norm, 0.0, 0, Sxy, D1, D2);
  mean_f32<<<blocks_for(D1), 256, 0, st>>>(sum_x, n_frames, mean_x, D1);
  mean_f32<<<blocks_for(D2), 256, 0, st>>>(sum_y, n_frames, mean_y, D2);
  TD_CHECK_CUDA(cudaGetLastError());
  const double one = 1.0, zero = 0.0;
  int hinfo[2] = {0, 0};
  // ---- fast path: Cholesky whitening -------------------------------------------------------------
  // With reg > eps_eig both covariances are positive definite and the reference's pseudo-inverse
  // square root keeps every eigenvalue (cca.py:349-360), so the canonical pairs are those of the
  // generalised problem and any whitening gives them (up to the joint sign of a pair):
  //   Sxx = L1 L1', Syy = L2 L2', T = L1^-1 Sxy L2^-T = U E V', rot_x = L1^-T U, rot_y = L2^-T V.
  // The SVD comes from the eigenvectors of the smaller Gram of T; it needs sigma_dim well above the
  // rounding floor of sigma_1^2, otherwise (or if a factorisation fails) the eigendecomposition
  // path below runs.
  if (reg > 10.0 * eps_eig) {
    int lw1 = 0, lw2 = 0;
    TD_CHECK_SOLVER(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, D1, Sxx, D1, &lw1));
    TD_CHECK_SOLVER(cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, D2, Syy, D2, &lw2));
    double* pwork;
    if ((rc = sc.get((size_t)std::max(lw1, lw2), &pwork))) return rc;
    TD_CHECK_SOLVER(cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, D1, Sxx, D1, pwork, lw1, info));
    TD_CHECK_SOLVER(cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, D2, Syy, D2, pwork, lw2, info + 1));
    // column-major Sxy buffer = Sxy' [D2, D1]:  T' = L2^-1 Sxy' L1^-T
    TD_CHECK_SOLVER(cublasDtrsm(h->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT,
                                D2, D1, &one, Syy, D2, Sxy, D2));
    TD_CHECK_SOLVER(cublasDtrsm(h->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT,
                                D2, D1, &one, Sxx, D1, Sxy, D2));
    const bool small_is_y = D2 <= D1;                  // Gram over the smaller side
    const int ns = small_is_y ? D2 : D1, nl = small_is_y ? D1 : D2;
    double *C, *ws, *top_s, *top_l, *sigma;
    if ((rc = sc.get((long long)ns * ns, &C)) || (rc = sc.get(ns, &ws)) || (rc = sc.get((long long)ns * dim, &top_s)) ||
        (rc = sc.get((long long)nl * dim, &top_l)) || (rc = sc.get(dim, &sigma)))
      return rc;
    if (small_is_y)   // C = T' T'^T  [D2, D2]
      TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, D2, D2, D1, &one, Sxy, D2, Sxy, D2, &zero, C, D2));
    else              // C = T'^T T'  [D1, D1]
      TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, D1, D1, D2, &one, Sxy, D2, Sxy, D2, &zero, C, D1));
    int* info3;
    if ((rc = sc.get(1, &info3))) return rc;
    if ((rc = eigh(h, sc, C, ns, ws, info3))) return rc;
    take_top_eigvecs<<<blocks_for((long long)ns * dim), 256, 0, st>>>(C, ws, ns, dim, top_s, sigma);
    if (small_is_y)   // top_s = V (y side, [D2, dim]); U = T V / sigma = T'^T V / sigma  [D1, dim]
      TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, D1, dim, D2, &one, Sxy, D2, top_s, D2, &zero, top_l, D1));
    else              // top_s = U (x side, [D1, dim]); V = T' U / sigma  [D2, dim]
      TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, D2, dim, D1, &one, Sxy, D2, top_s, D1, &zero, top_l, D2));
    scale_cols_recip<<<blocks_for((long long)nl * dim), 256, 0, st>>>(top_l, sigma, nl, dim);
    TD_CHECK_CUDA(cudaGetLastError());
    double* ux = small_is_y ? top_l : top_s;           // [D1, dim]
    double* vy = small_is_y ? top_s : top_l;           // [D2, dim]
    std::vector<double> hs(dim);
    int hinfo3[3] = {0, 0, 0};
    TD_CHECK_CUDA(cudaMemcpyAsync(hinfo3, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    TD_CHECK_CUDA(cudaMemcpyAsync(hinfo3 + 2, info3, sizeof(int), cudaMemcpyDeviceToHost, st));
    TD_CHECK_CUDA(cudaMemcpyAsync(hs.data(), sigma, dim * sizeof(double), cudaMemcpyDeviceToHost, st));
    TD_CHECK_CUDA(cudaStreamSynchronize(st));
    const bool usable = hinfo3[0] == 0 && hinfo3[1] == 0 && hinfo3[2] == 0 && hs[0] > 0.0 &&
                        hs[dim - 1] > 1e-4 * hs[0];
    if (usable) {
      // rot_x = L1^-T U, rot_y = L2^-T V
      TD_CHECK_SOLVER(cublasDtrsm(h->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT,
                                  D1, dim, &one, Sxx, D1, ux, D1));
      TD_CHECK_SOLVER(cublasDtrsm(h->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT,
                                  D2, dim, &one, Syy, D2, vy, D2));
      normalise_signs<<<dim, 256, 0, st>>>(ux, D1, vy, D2);
      to_rowmajor_f32<<<blocks_for((long long)D1 * dim), 256, 0, st>>>(ux, D1, rot_x, D1, dim);
      to_rowmajor_f32<<<blocks_for((long long)D2 * dim), 256, 0, st>>>(vy, D2, rot_y, D2, dim);
      copy_f32<<<blocks_for(dim), 256, 0, st>>>(sigma, e_out, dim);
      TD_CHECK_CUDA(cudaGetLastError());
      TD_CHECK_CUDA(cudaStreamSynchronize(st));
      return TD_OK;
    }
    // fall back: rebuild the covariances the factorisations overwrote
    make_cov<<<blocks_for(n11), 256, 0, st>>>(Gxx, sum_x, sum_x, n_frames, n_norm, reg, 1, Sxx, D1, D1);
    make_cov<<<blocks_for(n22), 256, 0, st>>>(Gyy, sum_y, sum_y, n_frames, n_norm, reg, 1, Syy, D2, D2);
    make_cov<<<blocks_for(n12), 256, 0, st>>>(Gxy, sum_x, sum_y, n_frames, n_norm, 0.0, 0, Sxy, D1, D2);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  if ((rc = eigh(h, sc, Sxx, D1, w1, info))) return rc;
  if ((rc = eigh(h, sc, Syy, D2, w2, info + 1))) return rc;
  TD_CHECK_CUDA(cudaMemcpyAsync(hinfo, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  if (hinfo[0] || hinfo[1]) { set_error("td_cca_solve: syevd did not converge (info %d, %d)", hinfo[0], hinfo[1]); return TD_ERR_SOLVER; }
  // K = V diag(w^-1/2) V' over eigenvalues > eps   (cca.py:349-360)
  scale_cols_inv4<<<blocks_for(n11), 256, 0, st>>>(Sxx, w1, eps_eig, D1);
  scale_cols_inv4<<<blocks_for(n22), 256, 0, st>>>(Syy, w2, eps_eig, D2);
  TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, D1, D1, D1, &one, Sxx, D1, Sxx, D1, &zero, K11, D1));
  TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, D2, D2, D2, &one, Syy, D2, Syy, D2, &zero, K22, D2));
  // Row-major Sxy [D1, D2] is column-major Sxy' [D2, D1].  T' = K22 Sxy' K11   (cca.py:361)
  TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, D2, D1, D2, &one, K22, D2, Sxy, D2, &zero, tmp, D2));
  TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, D2, D1, D1, &one, tmp, D2, K11, D1, &zero, Tt, D2));
  // SVD of the taller orientation M [m, n] (gesvd needs m >= n); M = L E R'.
  const bool tall_is_T = D1 > D2;           // M = T (D1 x D2)  else M = T' (D2 x D1)
  const int m = tall_is_T ? D1 : D2, nn = tall_is_T ? D2 : D1;
  double* M = Tt;
  if (tall_is_T) {
    TD_CHECK_SOLVER(cublasDgeam(h->blas, CUBLAS_OP_T, CUBLAS_OP_N, D1, D2, &one, Tt, D2, &zero, tmp, D1, tmp, D1));
    M = tmp;
  }
  double *Sv, *U, *VT, *work;
  int lwork = 0;
  if ((rc = sc.get(nn, &Sv)) || (rc = sc.get((long long)m * nn, &U)) || (rc = sc.get((long long)nn * nn, &VT))) return rc;
  TD_CHECK_SOLVER(cusolverDnDgesvd_bufferSize(h->solver, m, nn, &lwork));
  if ((rc = sc.get((size_t)lwork, &work))) return rc;
  TD_CHECK_SOLVER(cusolverDnDgesvd(h->solver, 'S', 'S', m, nn, M, m, Sv, U, m, VT, nn, work, lwork, nullptr, info));
  TD_CHECK_CUDA(cudaMemcpyAsync(hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  if (hinfo[0]) { set_error("td_cca_solve: gesvd failed (info %d)", hinfo[0]); return TD_ERR_SOLVER; }
  // Left vectors L = U[:, :dim] (column-major), right vectors R[:, :dim] = (VT[:dim, :])'.
  double *rx, *ry;
  if ((rc = sc.get((long long)D1 * dim, &rx)) || (rc = sc.get((long long)D2 * dim, &ry))) return rc;
  if (tall_is_T) {   // T = U E V': rot_x = K11 U, rot_y = K22 V     (cca.py:365-366)
    TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, D1, dim, D1, &one, K11, D1, U, m, &zero, rx, D1));
    TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, D2, dim, D2, &one, K22, D2, VT, nn, &zero, ry, D2));
  } else {           // T' = L E R' => T = R E L': rot_x = K11 R, rot_y = K22 L
    TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_T, D1, dim, D1, &one, K11, D1, VT, nn, &zero, rx, D1));
    TD_CHECK_SOLVER(cublasDgemm(h->blas, CUBLAS_OP_N, CUBLAS_OP_N, D2, dim, D2, &one, K22, D2, U, m, &zero, ry, D2));
  }
  normalise_signs<<<dim, 256, 0, st>>>(rx, D1, ry, D2);
  to_rowmajor_f32<<<blocks_for((long long)D1 * dim), 256, 0, st>>>(rx, D1, rot_x, D1, dim);
  to_rowmajor_f32<<<blocks_for((long long)D2 * dim), 256, 0, st>>>(ry, D2, rot_y, D2, dim);
  copy_f32<<<blocks_for(dim), 256, 0, st>>>(Sv, e_out, dim);
  TD_CHECK_CUDA(cudaGetLastError());
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  return TD_OK;
}
