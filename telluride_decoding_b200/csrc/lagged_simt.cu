// CUDA-core kernels that read the raw time-major streams directly through the lag mapping
//   X_s[t, k*C + c] = x_s[t - pre + k, c]  (0 outside the segment)      brain_data.py:446-455
//   * gram_simt:   fp32 FMA Gram, on-device cross-check of the tcgen05 kernel (method 1)
//   * thin:        X'[v | 1] and [v | 1]'[v | 1] in float64            brain_model.py:436-439
//   * predict:     (X - mean) W + b, the Dense / CCA-rotation forward   brain_model.py:340-341,
//                                                                       cca.py:157-158
//   * rows:        the lagged rows themselves (compat iterator only)
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace td {

namespace {

struct MultiView {
  StreamView s[2];
  int n_streams;
  int64_t width[2];     // channels * lags
  int64_t Dz;
};

// Element (row t of segment seg, design column j) of Z = [X_0 | X_1].
__device__ __forceinline__ float design_at(const MultiView& mv, int seg, int64_t t, int64_t j) {
  int q = 0;
  if (mv.n_streams == 2 && j >= mv.width[0]) { q = 1; j -= mv.width[0]; }
  const StreamView& s = mv.s[q];
  const int k = (int)(j / s.channels);
  const int c = (int)(j - (int64_t)k * s.channels);
  const int64_t u = t - s.pre + k;
  if (u < 0 || u >= s.seg_len[seg]) return 0.0f;
  return __ldg(s.data + (s.seg_start[seg] + u) * s.ld + c);
}

constexpr int GT = 64;   // output tile
constexpr int GK = 16;   // frames per step

// One block: 64x64 tile of the upper triangle, a slice of segments (blockIdx.z).
__global__ void __launch_bounds__(256) gram_simt_kernel(MultiView mv, const int64_t* seg_first,
                                                        const int64_t* seg_keep, int n_seg,
                                                        int seg_per_z, float* G, int64_t ldg) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  __shared__ float As[GK][GT + 1];
  __shared__ float Bs[GK][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  const int s0 = blockIdx.z * seg_per_z, s1 = min(n_seg, s0 + seg_per_z);
  for (int seg = s0; seg < s1; ++seg) {
    const int64_t keep = seg_keep[seg];
    for (int64_t t0 = seg_first ? seg_first[seg] : 0; t0 < keep; t0 += GK) {
      for (int e = threadIdx.x; e < GK * GT; e += 256) {
        const int r = e / GT, col = e % GT;
        const int64_t t = t0 + r;
        const int64_t ja = (int64_t)bi * GT + col, jb = (int64_t)bj * GT + col;
        As[r][col] = (t < keep && ja < mv.Dz) ? design_at(mv, seg, t, ja) : 0.0f;
        Bs[r][col] = (t < keep && jb < mv.Dz) ? design_at(mv, seg, t, jb) : 0.0f;
      }
      __syncthreads();
      float part[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) part[a][b] = 0.0f;
#pragma unroll
      for (int r = 0; r < GK; ++r) {
        float av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = As[r][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = Bs[r][tx * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) part[a][b] = fmaf(av[a], bv[b], part[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] += (double)part[a][b];
      __syncthreads();
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t i = (int64_t)bi * GT + ty * 4 + a, j = (int64_t)bj * GT + tx * 4 + b;
      if (i < mv.Dz && j < mv.Dz && j >= i) atomicAdd(G + i * ldg + j, (float)acc[a][b]);
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int THIN_TB = 256;      // frames per block
constexpr int THIN_CB = 64;       // channels per block
constexpr int THIN_QMAX = 16;
constexpr int THIN_KT = 8;        // lags per thread (register window)
constexpr int THIN_QFAST = 4;     // target columns handled by the register-tiled path
constexpr int THIN_SLICES = 64;   // partial-sum copies: blocks add into slice (block % 64) to keep
                                  // same-address float64 atomics from serialising, then one reduce

// Register-tiled core of thin_kernel for QT target columns: THIN_KT consecutive lags of one channel.
// The frames x[r+k0 .. r+k0+KT) slide through registers, so one shared-memory load of x and QT of the
// targets feed KT*QT FMAs; fp32 partial sums are carried into float64 every 32 frames.
template <int QT>
__device__ __forceinline__ void thin_fast(const float* __restrict__ xs, const float* __restrict__ vs, int rows,
                                          int halo_rows, int cl, int k0, int L, int channels, int c,
                                          double* __restrict__ xtv) {
  double accd[THIN_KT][QT];
#pragma unroll
  for (int i = 0; i < THIN_KT; ++i)
#pragma unroll
    for (int q = 0; q < QT; ++q) accd[i][q] = 0.0;
  for (int r0 = 0; r0 < rows; r0 += 32) {
    float accf[THIN_KT][QT];
#pragma unroll
    for (int i = 0; i < THIN_KT; ++i)
#pragma unroll
      for (int q = 0; q < QT; ++q) accf[i][q] = 0.0f;
    const int r1 = min(rows, r0 + 32);
    float win[THIN_KT];
#pragma unroll
    for (int i = 0; i < THIN_KT - 1; ++i)
      win[i + 1] = (r0 + k0 + i < halo_rows) ? xs[(r0 + k0 + i) * THIN_CB + cl] : 0.0f;
    for (int r = r0; r < r1; ++r) {
#pragma unroll
      for (int i = 0; i < THIN_KT - 1; ++i) win[i] = win[i + 1];
      // (loaded by frame, not by lag validity: the value slides down to the lower lags)
      win[THIN_KT - 1] = (r + k0 + THIN_KT - 1 < halo_rows) ? xs[(r + k0 + THIN_KT - 1) * THIN_CB + cl] : 0.0f;
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        const float vv = vs[r * QT + q];
#pragma unroll
        for (int i = 0; i < THIN_KT; ++i) accf[i][q] = fmaf(win[i], vv, accf[i][q]);
      }
    }
#pragma unroll
    for (int i = 0; i < THIN_KT; ++i)
#pragma unroll
      for (int q = 0; q < QT; ++q) accd[i][q] += (double)accf[i][q];
  }
#pragma unroll
  for (int i = 0; i < THIN_KT; ++i) {
    if (k0 + i >= L) continue;
    const int64_t j = (int64_t)(k0 + i) * channels + c;
#pragma unroll
    for (int q = 0; q < QT; ++q) atomicAdd(xtv + j * QT + q, accd[i][q]);
  }
}

// Block = (256-frame piece of a segment, 64-channel block).  The raw frames it needs (piece + lag
// halo) and the targets are staged in shared memory once; thread j = (lag k, channel c) then walks
// the piece with fp32 FMAs, carried into float64 every 32 frames.
__global__ void __launch_bounds__(256) thin_kernel(StreamView x, StreamView v, int has_v, int Q,
                                                   const int64_t* seg_keep, const int* blk_seg,
                                                   const int64_t* blk_t0, double* xtv_slices,
                                                   double* vtv, int per_segment,
                                                   const int* __restrict__ seg_slot) {
  extern __shared__ float thin_sm[];
  const int seg = blk_seg[blockIdx.x];
  const int64_t t0 = blk_t0[blockIdx.x];
  const int64_t keep = seg_keep[seg];
  const int rows = (int)min((int64_t)THIN_TB, keep - t0);
  const int L = x.lags;
  const int halo_rows = rows + L - 1;
  const int c0 = blockIdx.y * THIN_CB;
  float* xs = thin_sm;                                   // [halo_rows][THIN_CB]
  float* vs = thin_sm + (size_t)(THIN_TB + L - 1) * THIN_CB;   // [rows][Q]
  const int64_t len = x.seg_len[seg];
  const float* base = x.data + x.seg_start[seg] * x.ld;
  for (int e = threadIdx.x; e < halo_rows * THIN_CB; e += blockDim.x) {
    const int r = e / THIN_CB, cl = e % THIN_CB;
    const int64_t u = t0 + r - x.pre;
    const int c = c0 + cl;
    xs[e] = (u >= 0 && u < len && c < x.channels) ? __ldg(base + u * x.ld + c) : 0.0f;
  }
  for (int e = threadIdx.x; e < rows * Q; e += blockDim.x) {
    const int r = e / Q, q = e % Q;
    float val = 1.0f;
    if (q < Q - 1) {
      const int64_t u = t0 + r;
      val = (has_v && u < v.seg_len[seg]) ? __ldg(v.data + (v.seg_start[seg] + u) * v.ld + q) : 0.0f;
    }
    vs[e] = val;
  }
  __syncthreads();
  const int64_t Dx = (int64_t)x.channels * L;
  // One output per segment (leave-one-trial-out statistics), or 64 slices that a second kernel adds up.
  const int slot = seg_slot ? seg_slot[seg] : seg;
  double* xtv = per_segment ? xtv_slices + (size_t)slot * Dx * Q
                            : xtv_slices + (size_t)((blockIdx.x + blockIdx.y) % THIN_SLICES) * Dx * Q;
  if (per_segment) vtv += (size_t)slot * Q * Q;
  // Thread = (channel, group of THIN_KT consecutive lags): the frames x[r+k .. r+k+KT) slide
  // through registers, so one shared-memory load of x and Q of the targets feed KT*Q FMAs.
  const int n_kg = (L + THIN_KT - 1) / THIN_KT;
  for (int jj = threadIdx.x; jj < n_kg * THIN_CB; jj += blockDim.x) {
    const int kg = jj / THIN_CB, cl = jj % THIN_CB;
    const int c = c0 + cl;
    if (c >= x.channels) continue;
    const int k0 = kg * THIN_KT;
    if (Q <= THIN_QFAST) {
      switch (Q) {
        case 1: thin_fast<1>(xs, vs, rows, halo_rows, cl, k0, L, x.channels, c, xtv); break;
        case 2: thin_fast<2>(xs, vs, rows, halo_rows, cl, k0, L, x.channels, c, xtv); break;
        case 3: thin_fast<3>(xs, vs, rows, halo_rows, cl, k0, L, x.channels, c, xtv); break;
        default: thin_fast<4>(xs, vs, rows, halo_rows, cl, k0, L, x.channels, c, xtv); break;
      }
    } else {
      for (int i = 0; i < THIN_KT && k0 + i < L; ++i) {
        const int k = k0 + i;
        double accd[THIN_QMAX];
#pragma unroll
        for (int q = 0; q < THIN_QMAX; ++q) accd[q] = 0.0;
        for (int r0 = 0; r0 < rows; r0 += 32) {
          float accf[THIN_QMAX];
#pragma unroll
          for (int q = 0; q < THIN_QMAX; ++q) accf[q] = 0.0f;
          const int r1 = min(rows, r0 + 32);
          for (int r = r0; r < r1; ++r) {
            const float xv = xs[(r + k) * THIN_CB + cl];
#pragma unroll
            for (int q = 0; q < THIN_QMAX; ++q)
              if (q < Q) accf[q] = fmaf(xv, vs[r * Q + q], accf[q]);
          }
#pragma unroll
          for (int q = 0; q < THIN_QMAX; ++q) accd[q] += (double)accf[q];
        }
        const int64_t j = (int64_t)k * x.channels + c;
#pragma unroll
        for (int q = 0; q < THIN_QMAX; ++q)
          if (q < Q) atomicAdd(xtv + j * Q + q, accd[q]);
      }
    }
  }
  if (blockIdx.y == 0) {
    for (int e = threadIdx.x; e < Q * Q; e += blockDim.x) {
      const int a = e / Q, b = e % Q;
      double s = 0.0;
      for (int r = 0; r < rows; ++r) s += (double)vs[r * Q + a] * (double)vs[r * Q + b];
      atomicAdd(vtv + e, s);
    }
  }
}

__global__ void thin_reduce_kernel(const double* __restrict__ slices, double* __restrict__ out,
                                   int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < THIN_SLICES; ++k) s += slices[(size_t)k * n + i];
  out[i] = s;
}

// ---------------------------------------------------------------------------------------------
constexpr int PR_ROWS = 4;     // rows per warp

// PR_QC = output columns per pass (1, 2, 4 or 8: the smallest that covers Dy, so that a single-target
// decoder does not multiply seven columns of zeros)
template <int PR_QC>
__global__ void __launch_bounds__(256) predict_kernel(StreamView x, const int64_t* seg_keep,
                                                      const int64_t* out_row, const int* blk_seg,
                                                      const int64_t* blk_t0, int rows_per_block,
                                                      const float* __restrict__ W,
                                                      const float* __restrict__ bias,
                                                      const float* __restrict__ mean, int Dy,
                                                      float* out, int64_t ldo, int64_t w_seg_stride,
                                                      int64_t b_seg_stride) {
  const int seg = blk_seg[blockIdx.x];
  W += (int64_t)seg * w_seg_stride;            // one weight matrix per segment (leave-one-trial-out), or shared
  if (bias) bias += (int64_t)seg * b_seg_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t keep = seg_keep[seg];
  const int64_t t0 = blk_t0[blockIdx.x] + (int64_t)warp * PR_ROWS;
  if (t0 >= keep || warp * PR_ROWS >= rows_per_block) return;
  const int64_t len = x.seg_len[seg];
  const float* base = x.data + x.seg_start[seg] * x.ld;
  const int64_t Dx = (int64_t)x.channels * x.lags;
  for (int q0 = 0; q0 < Dy; q0 += PR_QC) {
    float acc[PR_ROWS][PR_QC];
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r)
#pragma unroll
      for (int q = 0; q < PR_QC; ++q) acc[r][q] = 0.0f;
    int k = lane / x.channels, c = lane % x.channels;
    for (int64_t j = lane; j < Dx; j += 32) {
      float w[PR_QC];
#pragma unroll
      for (int q = 0; q < PR_QC; ++q) w[q] = (q0 + q < Dy) ? __ldg(W + j * Dy + q0 + q) : 0.0f;
      const float mu = mean ? __ldg(mean + j) : 0.0f;
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) {
        const int64_t u = t0 + r - x.pre + k;
        float xv = (u >= 0 && u < len) ? __ldg(base + u * x.ld + c) : 0.0f;
        xv -= mu;
#pragma unroll
        for (int q = 0; q < PR_QC; ++q) acc[r][q] = fmaf(xv, w[q], acc[r][q]);
      }
      c += 32;
      while (c >= x.channels) { c -= x.channels; ++k; }
    }
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r)
#pragma unroll
      for (int q = 0; q < PR_QC; ++q) {
        float s = acc[r][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && q0 + q < Dy && t0 + r < keep)
          out[(out_row[seg] + t0 + r) * ldo + q0 + q] = s + (bias ? bias[q0 + q] : 0.0f);
      }
  }
}


// ---------------------------------------------------------------------------------------------
// Tiled forward: a block takes PT_ROWS consecutive design rows of one segment.  Per 32-channel block it
// stages the raw frames the rows need (rows + lag halo, zero outside the trial) and the weights of
// those channels in shared memory; thread r then owns rows r, r+128, r+256, r+384: every weight is
// read once for four rows, every x from shared memory (row stride 33 words: conflict free), so the
// kernel is bound by the one pass over the raw frames instead of L1 traffic per (row, column).
// (x - mean) W = x W - mean W: the constant is computed once per block.
constexpr int PT_THREADS = 128;
constexpr int PT_RPT = 4;                       // rows per thread
constexpr int PT_ROWS = PT_THREADS * PT_RPT;    // 512
constexpr int PT_CB = 32;                       // channels per pass

template <int QC>
__global__ void __launch_bounds__(PT_THREADS) predict_tiled_kernel(StreamView x, const int64_t* seg_keep,
                                                                   const int64_t* out_row, const int* blk_seg,
                                                                   const int64_t* blk_t0,
                                                                   const float* __restrict__ W,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ mean, int Dy,
                                                                   float* out, int64_t ldo, int64_t w_seg_stride,
                                                                   int64_t b_seg_stride) {
  extern __shared__ float pt_sm[];
  const int seg = blk_seg[blockIdx.x];
  const int64_t t0 = blk_t0[blockIdx.x];
  const int64_t keep = seg_keep[seg];
  const int64_t len = x.seg_len[seg];
  const int L = x.lags, C = x.channels;
  const int rows = (int)min((int64_t)PT_ROWS, keep - t0);
  const int nf = rows + L - 1;                                  // raw frames t0 - pre .. t0 - pre + nf
  float* xs = pt_sm;                                            // [PT_ROWS + L - 1][33]
  float* ws = pt_sm + (size_t)(PT_ROWS + L - 1) * (PT_CB + 1);  // [L][PT_CB][QC]
  __shared__ float s_off[QC];
  __shared__ float s_red[PT_THREADS / 32][QC];
  const float* base = x.data + x.seg_start[seg] * x.ld;
  W += (int64_t)seg * w_seg_stride;
  if (bias) bias += (int64_t)seg * b_seg_stride;
  const int64_t Dx = (int64_t)C * L;
  const int r = threadIdx.x;
  for (int q0 = 0; q0 < Dy; q0 += QC) {
    // constant -mean W of these output columns
    if (mean) {
      float part[QC];
#pragma unroll
      for (int q = 0; q < QC; ++q) part[q] = 0.0f;
      for (int64_t j = r; j < Dx; j += PT_THREADS) {
        const float mu = __ldg(mean + j);
#pragma unroll
        for (int q = 0; q < QC; ++q)
          if (q0 + q < Dy) part[q] = fmaf(mu, __ldg(W + j * Dy + q0 + q), part[q]);
      }
#pragma unroll
      for (int q = 0; q < QC; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[q] += __shfl_xor_sync(0xffffffffu, part[q], o);
        if ((r & 31) == 0) s_red[r >> 5][q] = part[q];
      }
      __syncthreads();
      if (r < QC) {
        float t = 0.0f;
        for (int w = 0; w < PT_THREADS / 32; ++w) t += s_red[w][r];
        s_off[r] = t;
      }
      __syncthreads();
    }
    float acc[PT_RPT][QC];
#pragma unroll
    for (int j = 0; j < PT_RPT; ++j)
#pragma unroll
      for (int q = 0; q < QC; ++q) acc[j][q] = 0.0f;
    for (int c0 = 0; c0 < C; c0 += PT_CB) {
      for (int e = r; e < nf * PT_CB; e += PT_THREADS) {
        const int f = e / PT_CB, cc = e % PT_CB;
        const int64_t u = t0 - x.pre + f;
        xs[f * (PT_CB + 1) + cc] = (u >= 0 && u < len && c0 + cc < C) ? __ldg(base + u * x.ld + c0 + cc) : 0.0f;
      }
      for (int e = r; e < L * PT_CB * QC; e += PT_THREADS) {
        const int q = e % QC, cc = (e / QC) % PT_CB, k = e / (QC * PT_CB);
        ws[e] = (c0 + cc < C && q0 + q < Dy) ? __ldg(W + ((int64_t)k * C + c0 + cc) * Dy + q0 + q) : 0.0f;
      }
      __syncthreads();
      for (int k = 0; k < L; ++k) {
        const float* xr = xs + (size_t)(r + k) * (PT_CB + 1);
        const float* wk = ws + (size_t)k * PT_CB * QC;
#pragma unroll 8
        for (int cc = 0; cc < PT_CB; ++cc) {
          float w[QC];
#pragma unroll
          for (int q = 0; q < QC; ++q) w[q] = wk[cc * QC + q];
#pragma unroll
          for (int j = 0; j < PT_RPT; ++j) {
            // rows beyond `rows` read frames of the tile that were never staged: they are not written out
            const float xv = (r + PT_THREADS * j < rows) ? xr[(size_t)PT_THREADS * j * (PT_CB + 1) + cc] : 0.0f;
#pragma unroll
            for (int q = 0; q < QC; ++q) acc[j][q] = fmaf(xv, w[q], acc[j][q]);
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < PT_RPT; ++j) {
      const int row = r + PT_THREADS * j;
      if (row < rows) {
#pragma unroll
        for (int q = 0; q < QC; ++q)
          if (q0 + q < Dy)
            out[(out_row[seg] + t0 + row) * ldo + q0 + q] =
                acc[j][q] + (bias ? bias[q0 + q] : 0.0f) - (mean ? s_off[q] : 0.0f);
      }
    }
    __syncthreads();
  }
}

__global__ void rows_kernel(StreamView x, const int64_t* seg_keep, const int64_t* out_row,
                            const int* blk_seg, const int64_t* blk_t0, int rows_per_block,
                            float* out, int64_t ldo) {
  const int seg = blk_seg[blockIdx.x];
  const int64_t keep = seg_keep[seg];
  const int64_t len = x.seg_len[seg];
  const int64_t Dx = (int64_t)x.channels * x.lags;
  const float* base = x.data + x.seg_start[seg] * x.ld;
  for (int r = 0; r < rows_per_block; ++r) {
    const int64_t t = blk_t0[blockIdx.x] + r;
    if (t >= keep) break;
    for (int64_t j = threadIdx.x; j < Dx; j += blockDim.x) {
      const int k = (int)(j / x.channels);
      const int c = (int)(j - (int64_t)k * x.channels);
      const int64_t u = t - x.pre + k;
      out[(out_row[seg] + t) * ldo + j] = (u >= 0 && u < len) ? __ldg(base + u * x.ld + c) : 0.0f;
    }
  }
}

// Ledoit-Wolf statistic (brain_model.py:440-443, :457-465).  The reference accumulates the Gram of the
// squared, running-mean-centred minibatch, z = (x - mean_b)^2, and then uses only the SUM of all its
// entries: sum_ij sum_t z_ti z_tj = sum_t (sum_i z_ti)^2.  One warp per design row computes the row
// sum over the implicit lagged design (lanes stride the columns), the block adds the squares.
__global__ void __launch_bounds__(256) lw_stat_kernel(StreamView x, const int64_t* seg_keep, const int* blk_seg,
                                                      const int64_t* blk_t0, int rows_per_block,
                                                      const int* __restrict__ seg_slot,
                                                      const float* __restrict__ mean, double* out) {
  __shared__ double part[8];
  const int seg = blk_seg[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t keep = seg_keep[seg];
  const int64_t len = x.seg_len[seg];
  const float* base = x.data + x.seg_start[seg] * x.ld;
  const int64_t Dx = (int64_t)x.channels * x.lags;
  const float* mu = mean + (int64_t)seg_slot[seg] * Dx;
  double acc = 0.0;
  for (int r = warp; r < rows_per_block; r += 8) {
    const int64_t t = blk_t0[blockIdx.x] + r;
    if (t >= keep) break;
    float s = 0.0f;
    int k = lane / x.channels, c = lane % x.channels;
    for (int64_t j = lane; j < Dx; j += 32) {
      const int64_t u = t - x.pre + k;
      const float xv = (u >= 0 && u < len) ? __ldg(base + u * x.ld + c) : 0.0f;
      const float dlt = xv - __ldg(mu + j);
      s = fmaf(dlt, dlt, s);
      c += 32;
      while (c >= x.channels) { c -= x.channels; ++k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    acc += (double)s * (double)s;
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += part[w];
    atomicAdd(out, tot);
  }
}

// Device copies of the segment tables of one stream (+ keep, out_row), stream-ordered.
struct DevTables {
  std::vector<void*> ptrs;
  cudaStream_t st;
  explicit DevTables(cudaStream_t s) : st(s) {}
  ~DevTables() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  template <typename T>
  int put(const T* host, size_t n, const T** dev) {
    std::vector<T> h(host, host + n);
    T* d = nullptr;
    int rc = upload(h, &d, st);
    if (rc) return rc;
    ptrs.push_back(d);
    *dev = d;
    return TD_OK;
  }
};

int make_view(const td_stream* s, int n_seg, DevTables& tabs, StreamView* v) {
  v->data = s->data_dev;
  v->ld = s->ld;
  v->channels = s->channels;
  v->pre = s->pre;
  v->lags = lags_of(*s);
  int rc = tabs.put(s->seg_start, n_seg, &v->seg_start);
  if (rc) return rc;
  return tabs.put(s->seg_len, n_seg, &v->seg_len);
}

// (segment, first row) of every block when each block covers `rows` kept rows of one segment.
int make_blocks(const int64_t* seg_keep, int n_seg, int rows, DevTables& tabs, const int** blk_seg,
                const int64_t** blk_t0, int64_t* n_blocks, const int64_t* seg_first = nullptr) {
  std::vector<int> bs;
  std::vector<int64_t> bt;
  for (int s = 0; s < n_seg; ++s)
    for (int64_t t = seg_first ? seg_first[s] : 0; t < seg_keep[s]; t += rows) { bs.push_back(s); bt.push_back(t); }
  *n_blocks = (int64_t)bs.size();
  int rc = tabs.put(bs.data(), bs.size(), blk_seg);
  if (rc) return rc;
  return tabs.put(bt.data(), bt.size(), blk_t0);
}

}  // namespace

int check_stream(const td_stream* s, const int64_t* seg_keep, int n_seg, const char* what) {
  TD_REQUIRE(s != nullptr, "%s: null stream", what);
  TD_REQUIRE(s->data_dev != nullptr, "%s: stream has no data", what);
  TD_REQUIRE(s->channels > 0, "%s: stream must have at least one channel, not %d", what, s->channels);
  TD_REQUIRE(s->pre >= 0, "%s: pre_context must be >= 0", what);
  TD_REQUIRE(s->post >= 0, "%s: post_context must be >= 0", what);
  TD_REQUIRE(s->ld >= s->channels, "%s: row stride %lld < channels %d", what, (long long)s->ld, s->channels);
  TD_REQUIRE(s->seg_start && s->seg_len, "%s: missing segment table", what);
  for (int i = 0; i < n_seg; ++i) {
    TD_REQUIRE(s->seg_start[i] >= 0 && s->seg_len[i] >= 0, "%s: negative segment bounds at %d", what, i);
    if (seg_keep)
      TD_REQUIRE(seg_keep[i] >= 0, "%s: negative keep count at segment %d", what, i);
  }
  return TD_OK;
}

int launch_gram_simt(const td_stream* streams, int n_streams, const int64_t* seg_first,
                     const int64_t* seg_keep, int n_seg, float* G, int64_t ldg, cudaStream_t st) {
  DevTables tabs(st);
  MultiView mv;
  mv.n_streams = n_streams;
  mv.Dz = 0;
  mv.width[0] = mv.width[1] = 0;
  for (int q = 0; q < n_streams; ++q) {
    int rc = make_view(&streams[q], n_seg, tabs, &mv.s[q]);
    if (rc) return rc;
    mv.width[q] = width_of(streams[q]);
    mv.Dz += mv.width[q];
  }
  if (n_streams == 1) mv.s[1] = mv.s[0];
  const int64_t* keep_dev;
  const int64_t* first_dev = nullptr;
  int rc = tabs.put(seg_keep, n_seg, &keep_dev);
  if (rc) return rc;
  if (seg_first && (rc = tabs.put(seg_first, n_seg, &first_dev))) return rc;
  TD_CHECK_CUDA(cudaMemset2DAsync(G, (size_t)ldg * sizeof(float), 0, (size_t)mv.Dz * sizeof(float), (size_t)mv.Dz, st));
  const int nb = (int)ceil_div(mv.Dz, GT);
  int zsplit = std::max(1, std::min(n_seg, (int)(1024 / std::max(1, nb * (nb + 1) / 2))));
  const int seg_per_z = (int)ceil_div(n_seg, zsplit);
  zsplit = (int)ceil_div(n_seg, seg_per_z);
  gram_simt_kernel<<<dim3(nb, nb, zsplit), 256, 0, st>>>(mv, first_dev, keep_dev, n_seg, seg_per_z, G, ldg);
  TD_CHECK_CUDA(cudaGetLastError());
  return td_symmetrize_f32(G, mv.Dz, ldg, 1, st);
}

}  // namespace td

using namespace td;

extern "C" int td_lagged_gram_range_f32(const td_stream* streams, int n_streams, const int64_t* seg_first,
                                        const int64_t* seg_keep, int n_seg, float* G, int64_t ldg,
                                        void* stream) {
  TD_REQUIRE(streams && seg_keep && G && n_seg > 0, "td_lagged_gram_range_f32: null argument");
  TD_REQUIRE(n_streams == 1 || n_streams == 2, "td_lagged_gram_range_f32: n_streams must be 1 or 2");
  int64_t Dz = 0;
  for (int q = 0; q < n_streams; ++q) {
    int rc = check_stream(&streams[q], seg_keep, n_seg, "td_lagged_gram_range_f32");
    if (rc) return rc;
    Dz += width_of(streams[q]);
  }
  TD_REQUIRE(ldg >= Dz, "td_lagged_gram_range_f32: ldg too small");
  return launch_gram_simt(streams, n_streams, seg_first, seg_keep, n_seg, G, ldg, (cudaStream_t)stream);
}

static int thin_impl(const td_stream* x, const td_stream* v, const int64_t* seg_first,
                     const int64_t* seg_keep, int n_seg, double* xtv, double* vtv, void* stream,
                     int per_segment, const int32_t* seg_slot = nullptr, int n_slots = 0);

extern "C" int td_lagged_thin_f64(const td_stream* x, const td_stream* v, const int64_t* seg_first,
                                  const int64_t* seg_keep, int n_seg, double* xtv, double* vtv,
                                  void* stream) {
  return thin_impl(x, v, seg_first, seg_keep, n_seg, xtv, vtv, stream, 0);
}

extern "C" int td_lagged_thin_per_segment_f64(const td_stream* x, const td_stream* v,
                                              const int64_t* seg_first, const int64_t* seg_keep,
                                              int n_seg, double* xtv, double* vtv, void* stream) {
  return thin_impl(x, v, seg_first, seg_keep, n_seg, xtv, vtv, stream, 1);
}

// Column sums of the lagged design per OUTPUT SLOT: segment i (rows [first, keep)) adds into slot
// seg_slot[i].  With the pieces of the minibatches as segments this gives `np.sum(x, axis=0)` of every
// minibatch (brain_model.py:438) in one launch.
extern "C" int td_lagged_column_sums_f64(const td_stream* x, const int64_t* seg_first, const int64_t* seg_keep,
                                         const int32_t* seg_slot, int n_seg, int n_slots, double* sums_dev,
                                         double* rows_dev, void* stream) {
  TD_REQUIRE(seg_slot && n_slots > 0 && sums_dev && rows_dev, "td_lagged_column_sums_f64: null argument");
  for (int i = 0; i < n_seg; ++i)
    TD_REQUIRE(seg_slot[i] >= 0 && seg_slot[i] < n_slots, "td_lagged_column_sums_f64: slot %d of segment %d out of range", seg_slot[i], i);
  return thin_impl(x, nullptr, seg_first, seg_keep, n_seg, sums_dev, rows_dev, stream, 1, seg_slot, n_slots);
}

static int thin_impl(const td_stream* x, const td_stream* v, const int64_t* seg_first,
                     const int64_t* seg_keep, int n_seg, double* xtv, double* vtv, void* stream,
                     int per_segment, const int32_t* seg_slot, int n_slots) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(x && seg_keep && xtv && vtv && n_seg > 0, "td_lagged_thin_f64: null argument");
  int rc = check_stream(x, seg_keep, n_seg, "td_lagged_thin_f64");
  if (rc) return rc;
  int Q = 1;
  if (v) {
    rc = check_stream(v, seg_keep, n_seg, "td_lagged_thin_f64");
    if (rc) return rc;
    TD_REQUIRE(v->pre == 0 && v->post == 0, "td_lagged_thin_f64: the thin stream must not have context");
    Q = v->channels + 1;
  }
  if (Q > THIN_QMAX) {
    set_error("td_lagged_thin_f64: %d target columns (+1) exceed %d; use td_lagged_gram_f32 with two streams", Q - 1, THIN_QMAX);
    return TD_ERR_UNSUPPORTED;
  }
  DevTables tabs(st);
  StreamView xv, vv;
  rc = make_view(x, n_seg, tabs, &xv);
  if (rc) return rc;
  vv = xv;
  if (v) { rc = make_view(v, n_seg, tabs, &vv); if (rc) return rc; }
  const int64_t* keep_dev;
  rc = tabs.put(seg_keep, n_seg, &keep_dev);
  if (rc) return rc;
  const int* blk_seg; const int64_t* blk_t0; int64_t nblk;
  rc = make_blocks(seg_keep, n_seg, THIN_TB, tabs, &blk_seg, &blk_t0, &nblk, seg_first);
  if (rc) return rc;
  const int64_t Dx = width_of(*x);
  const size_t n_out = per_segment ? (size_t)(seg_slot ? n_slots : n_seg) : 1;
  const int* slot_dev = nullptr;
  if (seg_slot && (rc = tabs.put(seg_slot, (size_t)n_seg, &slot_dev))) return rc;
  TD_CHECK_CUDA(cudaMemsetAsync(xtv, 0, n_out * Dx * Q * sizeof(double), st));
  TD_CHECK_CUDA(cudaMemsetAsync(vtv, 0, n_out * Q * Q * sizeof(double), st));
  if (nblk > 0) {
    const int L = lags_of(*x);
    const size_t smem = ((size_t)(THIN_TB + L - 1) * THIN_CB + (size_t)THIN_TB * Q) * sizeof(float);
    if (smem > 220 * 1024) {
      set_error("td_lagged_thin_f64: %d lags need %zu bytes of shared memory", L, smem);
      return TD_ERR_UNSUPPORTED;
    }
    // Partial-sum slices: owned by a guard so that every error path below frees them.
    struct Slices {
      double* p = nullptr;
      cudaStream_t st;
      explicit Slices(cudaStream_t s) : st(s) {}
      ~Slices() { if (p) cudaFreeAsync(p, st); }
    } slices(st);
    TD_CHECK_CUDA(cudaFuncSetAttribute(thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)nblk, (unsigned)ceil_div(x->channels, THIN_CB));
    if (per_segment) {
      thin_kernel<<<grid, 256, smem, st>>>(xv, vv, v ? 1 : 0, Q, keep_dev, blk_seg, blk_t0, xtv, vtv, 1, slot_dev);
      TD_CHECK_CUDA(cudaGetLastError());
    } else {
      const size_t slice_bytes = (size_t)THIN_SLICES * Dx * Q * sizeof(double);
      TD_CHECK_CUDA(cudaMallocAsync((void**)&slices.p, slice_bytes, st));
      TD_CHECK_CUDA(cudaMemsetAsync(slices.p, 0, slice_bytes, st));
      thin_kernel<<<grid, 256, smem, st>>>(xv, vv, v ? 1 : 0, Q, keep_dev, blk_seg, blk_t0, slices.p, vtv, 0, nullptr);
      TD_CHECK_CUDA(cudaGetLastError());
      thin_reduce_kernel<<<(unsigned)ceil_div(Dx * Q, 256), 256, 0, st>>>(slices.p, xtv, Dx * Q);
      TD_CHECK_CUDA(cudaGetLastError());
    }
  }
  return TD_OK;
}

namespace td {
namespace {
int predict_impl(const td_stream* x, const int64_t* seg_keep, const int64_t* out_row, int n_seg, const float* W,
                 const float* bias, const float* mean, int Dy, float* out, int64_t ldo, void* stream,
                 int64_t w_seg_stride, int64_t b_seg_stride);
}
}

extern "C" int td_lagged_predict_f32(const td_stream* x, const int64_t* seg_keep,
                                     const int64_t* out_row, int n_seg, const float* W,
                                     const float* bias, const float* mean, int Dy, float* out,
                                     int64_t ldo, void* stream) {
  return predict_impl(x, seg_keep, out_row, n_seg, W, bias, mean, Dy, out, ldo, stream, 0, 0);
}

extern "C" int td_lagged_predict_per_segment_f32(const td_stream* x, const int64_t* seg_keep,
                                                 const int64_t* out_row, int n_seg, const float* W,
                                                 const float* bias, int Dy, float* out, int64_t ldo,
                                                 void* stream) {
  TD_REQUIRE(x != nullptr, "td_lagged_predict_per_segment_f32: null argument");
  return predict_impl(x, seg_keep, out_row, n_seg, W, bias, nullptr, Dy, out, ldo, stream,
                      width_of(*x) * (int64_t)Dy, Dy);
}

namespace td {
namespace {
int predict_impl(const td_stream* x, const int64_t* seg_keep, const int64_t* out_row, int n_seg, const float* W,
                 const float* bias, const float* mean, int Dy, float* out, int64_t ldo, void* stream,
                 int64_t w_seg_stride, int64_t b_seg_stride) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(x && seg_keep && out_row && W && out && n_seg > 0, "td_lagged_predict_f32: null argument");
  TD_REQUIRE(Dy >= 1 && Dy <= 64 && ldo >= Dy, "td_lagged_predict_f32: Dy=%d ldo=%lld out of range", Dy, (long long)ldo);
  int rc = check_stream(x, seg_keep, n_seg, "td_lagged_predict_f32");
  if (rc) return rc;
  DevTables tabs(st);
  StreamView xv;
  rc = make_view(x, n_seg, tabs, &xv);
  if (rc) return rc;
  const int64_t *keep_dev, *orow_dev;
  rc = tabs.put(seg_keep, n_seg, &keep_dev); if (rc) return rc;
  rc = tabs.put(out_row, n_seg, &orow_dev); if (rc) return rc;
  // Tiled kernel whenever its tile (raw rows + halo, weights of 32 channels) fits shared memory.
  const int L = lags_of(*x);
  const int qc = (Dy == 1) ? 1 : (Dy == 2 ? 2 : 4);
  const size_t tiled_smem = ((size_t)(PT_ROWS + L - 1) * (PT_CB + 1) + (size_t)L * PT_CB * qc) * sizeof(float);
  if (tiled_smem <= 200 * 1024) {
    const int* blk_seg; const int64_t* blk_t0; int64_t nblk;
    rc = make_blocks(seg_keep, n_seg, PT_ROWS, tabs, &blk_seg, &blk_t0, &nblk);
    if (rc) return rc;
    if (nblk > 0) {
      if (qc == 1) {
        TD_CHECK_CUDA(cudaFuncSetAttribute(predict_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
        predict_tiled_kernel<1><<<(unsigned)nblk, PT_THREADS, tiled_smem, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
      } else if (qc == 2) {
        TD_CHECK_CUDA(cudaFuncSetAttribute(predict_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
        predict_tiled_kernel<2><<<(unsigned)nblk, PT_THREADS, tiled_smem, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
      } else {
        TD_CHECK_CUDA(cudaFuncSetAttribute(predict_tiled_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
        predict_tiled_kernel<4><<<(unsigned)nblk, PT_THREADS, tiled_smem, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
      }
      TD_CHECK_CUDA(cudaGetLastError());
    }
    return TD_OK;
  }
  const int rows_per_block = 8 * PR_ROWS;
  const int* blk_seg; const int64_t* blk_t0; int64_t nblk;
  rc = make_blocks(seg_keep, n_seg, rows_per_block, tabs, &blk_seg, &blk_t0, &nblk);
  if (rc) return rc;
  if (nblk > 0) {
    if (Dy == 1)
      predict_kernel<1><<<(unsigned)nblk, 256, 0, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, rows_per_block, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
    else if (Dy == 2)
      predict_kernel<2><<<(unsigned)nblk, 256, 0, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, rows_per_block, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
    else if (Dy <= 4)
      predict_kernel<4><<<(unsigned)nblk, 256, 0, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, rows_per_block, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
    else
      predict_kernel<8><<<(unsigned)nblk, 256, 0, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, rows_per_block, W, bias, mean, Dy, out, ldo, w_seg_stride, b_seg_stride);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  return TD_OK;
}
}  // namespace
}  // namespace td

extern "C" int td_lagged_rows_f32(const td_stream* x, const int64_t* seg_keep, const int64_t* out_row,
                                  int n_seg, float* out, int64_t ldo, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(x && seg_keep && out_row && out && n_seg > 0, "td_lagged_rows_f32: null argument");
  int rc = check_stream(x, seg_keep, n_seg, "td_lagged_rows_f32");
  if (rc) return rc;
  TD_REQUIRE(ldo >= width_of(*x), "td_lagged_rows_f32: ldo too small");
  DevTables tabs(st);
  StreamView xv;
  rc = make_view(x, n_seg, tabs, &xv);
  if (rc) return rc;
  const int64_t *keep_dev, *orow_dev;
  rc = tabs.put(seg_keep, n_seg, &keep_dev); if (rc) return rc;
  rc = tabs.put(out_row, n_seg, &orow_dev); if (rc) return rc;
  const int rows_per_block = 16;
  const int* blk_seg; const int64_t* blk_t0; int64_t nblk;
  rc = make_blocks(seg_keep, n_seg, rows_per_block, tabs, &blk_seg, &blk_t0, &nblk);
  if (rc) return rc;
  if (nblk > 0) {
    rows_kernel<<<(unsigned)nblk, 256, 0, st>>>(xv, keep_dev, orow_dev, blk_seg, blk_t0, rows_per_block, out, ldo);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  return TD_OK;
}

extern "C" int td_lagged_lw_stat_f64(const td_stream* x, const int64_t* seg_first, const int64_t* seg_keep,
                                     const int32_t* seg_slot, int n_seg, int n_slots, const float* mean_dev,
                                     double* out_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(x && seg_keep && seg_slot && mean_dev && out_dev && n_seg > 0 && n_slots > 0,
             "td_lagged_lw_stat_f64: null argument");
  int rc = check_stream(x, seg_keep, n_seg, "td_lagged_lw_stat_f64");
  if (rc) return rc;
  for (int i = 0; i < n_seg; ++i)
    TD_REQUIRE(seg_slot[i] >= 0 && seg_slot[i] < n_slots, "td_lagged_lw_stat_f64: slot %d of segment %d out of range", seg_slot[i], i);
  DevTables tabs(st);
  StreamView xv;
  rc = make_view(x, n_seg, tabs, &xv);
  if (rc) return rc;
  const int64_t* keep_dev;
  const int* slot_dev;
  rc = tabs.put(seg_keep, n_seg, &keep_dev); if (rc) return rc;
  rc = tabs.put(seg_slot, (size_t)n_seg, &slot_dev); if (rc) return rc;
  const int rows_per_block = 32;
  const int* blk_seg; const int64_t* blk_t0; int64_t nblk;
  rc = make_blocks(seg_keep, n_seg, rows_per_block, tabs, &blk_seg, &blk_t0, &nblk, seg_first);
  if (rc) return rc;
  TD_CHECK_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), st));
  if (nblk > 0) {
    lw_stat_kernel<<<(unsigned)nblk, 256, 0, st>>>(xv, keep_dev, blk_seg, blk_t0, rows_per_block, slot_dev, mean_dev, out_dev);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  return TD_OK;
}

// out[i] = sum_b in[b * stride + i] in float64 (the total Gram from the per-trial Grams: one coalesced
// pass over the batch instead of a strided library reduction).
namespace td { namespace {
__global__ void sum_batches_kernel(const float* __restrict__ in, long long n, long long stride, int batches,
                                   double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s0 = 0.0, s1 = 0.0;
  int b = 0;
  for (; b + 1 < batches; b += 2) {
    s0 += (double)__ldg(in + (long long)b * stride + i);
    s1 += (double)__ldg(in + (long long)(b + 1) * stride + i);
  }
  if (b < batches) s0 += (double)__ldg(in + (long long)b * stride + i);
  out[i] = s0 + s1;
}
} }

extern "C" int td_sum_batches_f32(const float* in_dev, int64_t n, int64_t stride, int batches, double* out_dev,
                                  void* stream) {
  TD_REQUIRE(in_dev && out_dev && n >= 0 && batches >= 1 && stride >= n, "td_sum_batches_f32: bad arguments");
  if (n == 0) return TD_OK;
  td::sum_batches_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(in_dev, n, stride, batches, out_dev);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}
