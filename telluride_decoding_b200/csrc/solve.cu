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
#include <vector>

#include "common.cuh"

namespace td {
namespace {

constexpr int MAX_LANES = 4;    // concurrent factorisations of a lambda sweep

struct Handles {
  cusolverDnHandle_t solver = nullptr;
  cublasHandle_t blas = nullptr;
  // side streams (+ their own cuSOLVER handles) so that the independent Cholesky factorisations of
  // a regularisation sweep overlap: one n ~ 2000 potrf cannot fill 148 SMs
  cudaStream_t lane_stream[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  cusolverDnHandle_t lane_solver[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_ready = nullptr;
  cudaEvent_t ev_done[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
};

int get_handles(Handles** out) {
  static Handles per_device[64];
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
    for (int i = 0; i < MAX_LANES; ++i) {
      TD_CHECK_CUDA(cudaStreamCreateWithFlags(&h.lane_stream[i], cudaStreamNonBlocking));
      TD_CHECK_CUDA(cudaEventCreateWithFlags(&h.ev_done[i], cudaEventDisableTiming));
      if (cusolverDnCreate(&h.lane_solver[i]) != CUSOLVER_STATUS_SUCCESS ||
          cusolverDnSetStream(h.lane_solver[i], h.lane_stream[i]) != CUSOLVER_STATUS_SUCCESS) {
        set_error("cusolverDnCreate (lane) failed");
        return TD_ERR_SOLVER;
      }
    }
  }
  *out = &h;
  return TD_OK;
}

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

}  // namespace
}  // namespace td

using namespace td;

extern "C" int td_ridge_solve(const double* G_aug, const double* C_aug, int Da, int Dy, double n,
                              const double* lambdas, int n_lambda, int method, float* W_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(G_aug && C_aug && lambdas && W_out, "td_ridge_solve: null argument");
  TD_REQUIRE(Da >= 1 && Dy >= 1 && n_lambda >= 1 && n > 0, "td_ridge_solve: Da=%d Dy=%d n_lambda=%d n=%g", Da, Dy, n_lambda, n);
  TD_REQUIRE(method == 0 || method == 1, "td_ridge_solve: unknown method %d", method);
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
  if ((rc = eigh(h, sc, Sxx, D1, w1, info))) return rc;
  if ((rc = eigh(h, sc, Syy, D2, w2, info + 1))) return rc;
  int hinfo[2] = {0, 0};
  TD_CHECK_CUDA(cudaMemcpyAsync(hinfo, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  TD_CHECK_CUDA(cudaStreamSynchronize(st));
  if (hinfo[0] || hinfo[1]) { set_error("td_cca_solve: syevd did not converge (info %d, %d)", hinfo[0], hinfo[1]); return TD_ERR_SOLVER; }
  // K = V diag(w^-1/2) V' over eigenvalues > eps   (cca.py:349-360)
  scale_cols_inv4<<<blocks_for(n11), 256, 0, st>>>(Sxx, w1, eps_eig, D1);
  scale_cols_inv4<<<blocks_for(n22), 256, 0, st>>>(Syy, w2, eps_eig, D2);
  const double one = 1.0, zero = 0.0;
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
