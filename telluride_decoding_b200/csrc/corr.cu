// Correlation, windowing and decision kernels (memory-bound; one pass over the signals).
//
//   td_column_moments_f64  -- Decoder.add_data_correlator sums      infer_decoder.py:299-303
//   td_windowed_corr_f32   -- per-frame correlation signal, reduction and sliding-window means
//                             infer_decoder.py:327-328, :441-450, result_store.py:262-271,
//                             infer.py:260-266; and per-window Pearson r brain_model.py:62-79
//   td_wta_decide_u8       -- AttentionDecoder.attention             attention_decoder.py:134
//
// Windows [i*S, i*S+W) of every requested shape start and end on multiples of
// g = gcd(all W, all S), so each segment is reduced ONCE into per-g-frame partial sums
// (coalesced loads, warp-shuffle reduction, float64) held in shared memory, and every window of
// every shape is then a short sum of consecutive partials: the data is read exactly once.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace td {
namespace {

constexpr int MAX_SIZES = 16;
constexpr int CORR_THREADS = 256;

struct WinSpec {
  int n_sizes;
  long long W[MAX_SIZES];
  long long S[MAX_SIZES];
  long long g;
};

// Stream-ordered device buffers freed on every exit path.
struct DevBufs {
  std::vector<void*> ptrs;
  cudaStream_t st;
  explicit DevBufs(cudaStream_t s) : st(s) {}
  ~DevBufs() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  template <typename T>
  int put(const std::vector<T>& host, T** dev) {
    int rc = upload(host, dev, st);
    if (rc == TD_OK) ptrs.push_back(*dev);
    return rc;
  }
  template <typename T>
  int alloc(size_t n, T** dev) {
    int rc = ensure_mempool_config();
    if (rc) return rc;
    TD_CHECK_CUDA(cudaMallocAsync((void**)dev, n * sizeof(T), st));
    ptrs.push_back(*dev);
    return TD_OK;
  }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-frame scalar of mode 0 / mode 1 (float32 arithmetic in the reference's operation order).
// reduction 5 ('lda'): slope * sum_d c[d] * w[d] + intercept in float64 (ScaledLDA.transform,
// scaled_lda.py:197-212, :345-355, column 0); lda = [w_0 .. w_{dims-1}, slope, intercept].
__device__ __forceinline__ double frame_signal(const float* a, const float* b, int dims, int mode,
                                               int reduction, const float* mu_a, const float* mu_b,
                                               const float* power, const double* lda) {
  if (mode == 0) return (double)a[0];
  if (reduction == 0 || reduction == 1) {
    const int d = reduction;
    return (double)(((a[d] - mu_a[d]) * (b[d] - mu_b[d])) / power[d]);
  }
  if (reduction == 5) {
    double acc = 0.0;
    for (int d = 0; d < dims; ++d)
      acc += (double)(((a[d] - mu_a[d]) * (b[d] - mu_b[d])) / power[d]) * lda[d];
    return lda[dims] * acc + lda[dims + 1];
  }
  // mean (2) and mean-squared (3): numpy float32 mean over the (contiguous) feature axis.
  // numpy sums < 8 values left to right and 8..128 values with eight interleaved accumulators
  // combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential remainder; same here.
  float r[8];
  float s = 0.0f;
  const int body = (dims >= 8) ? dims - (dims & 7) : 0;
  for (int d = 0; d < dims; ++d) {
    float c = ((a[d] - mu_a[d]) * (b[d] - mu_b[d])) / power[d];
    if (reduction == 3) {
      const float sg = (c > 0.0f) ? 1.0f : ((c < 0.0f) ? -1.0f : 0.0f);
      c = sg * (c * c);
    }
    if (d < body) {
      if (d < 8) r[d] = c; else r[d & 7] += c;
      if (d == body - 1) s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    } else {
      s += c;
    }
  }
  return (double)(s / (float)dims);
}

// a / p, correctly rounded like the reference's float32 division, from a reciprocal computed once
// per column: q = a*r, rem = a - q*p (exact in an FMA), q' = q + rem*r is RN(a/p) when r = RN(1/p)
// and nothing leaves the normal range (Markstein's final step).  The callers sum |q'| over the
// columns of a frame and redo the frame with real divisions when that sum is not below 1e30 (an
// overflow, an infinity or a NaN -- including the NaN stored as r for a power outside
// [1e-30, 1e30]); results below 1e-30 may differ from the division in their last bits, which is
// 1e-18 of anything a float64 window mean of O(1) signals can see.
__device__ __forceinline__ float div_fast(float a, float p, float r) {
  const float q = a * r;
  return fmaf(fmaf(-q, p, a), r, q);
}

// Same per-frame scalar from rows already held in registers (float4-loaded), DIMS known at compile
// time.  mu / power / lda point to shared memory.
template <int DIMS>
__device__ __forceinline__ double frame_signal_regs(const float (&a)[DIMS], const float (&b)[DIMS],
                                                    int reduction, const float* mu_a,
                                                    const float* mu_b, const float* power,
                                                    const float* rcp, const double* lda) {
  float c[DIMS];
  float mag = 0.0f;
#pragma unroll
  for (int d = 0; d < DIMS; ++d) {
    c[d] = div_fast((a[d] - mu_a[d]) * (b[d] - mu_b[d]), power[d], rcp[d]);
    mag += fabsf(c[d]);
  }
  if (!(mag < 1e30f)) {
#pragma unroll
    for (int d = 0; d < DIMS; ++d) c[d] = ((a[d] - mu_a[d]) * (b[d] - mu_b[d])) / power[d];
  }
  if (reduction == 0) return (double)c[0];
  if (reduction == 1) return (double)c[1];
  if (reduction == 5) {
    double acc = 0.0;
#pragma unroll
    for (int d = 0; d < DIMS; ++d) acc += (double)c[d] * lda[d];
    return lda[DIMS] * acc + lda[DIMS + 1];
  }
  if (reduction == 3) {
#pragma unroll
    for (int d = 0; d < DIMS; ++d) {
      const float sg = (c[d] > 0.0f) ? 1.0f : ((c[d] < 0.0f) ? -1.0f : 0.0f);
      c[d] = sg * (c[d] * c[d]);
    }
  }
  float s;
  if (DIMS < 8) {                      // numpy: plain left-to-right sum below 8 values
    s = 0.0f;
#pragma unroll
    for (int d = 0; d < DIMS; ++d) s += c[d];
  } else {                             // numpy: eight interleaved accumulators, pairwise combine
    float r[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) r[d] = c[d];
#pragma unroll
    for (int d = 8; d < DIMS; ++d) r[d & 7] += c[d];
    s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  }
  return (double)(s / (float)DIMS);
}

// ---------------------------------------------------------------------------------------------
// Phase 1: per-chunk partial sums (g = gcd of all widths and steps frames per chunk), written to
// global memory: part[(chunk_base[seg] + ch) * ncol + col], float64.
//   ncol: mode 0/1: 1 (the reduced signal)   mode 3: dims   mode 2: 5*dims (a, b, aa, bb, ab)
// A block handles CHUNKS_PER_BLOCK consecutive chunks of ONE segment (table blk_seg / blk_ch0):
// with g >= 32 one warp per chunk (coalesced 16-byte loads, two frames in flight per lane, float32
// per-frame arithmetic in the reference's order, float64 carry, shuffle reduction); with g < 32 one
// thread per chunk.  Partials are tiny next to the signals (8 bytes per g frames and column).
// ---------------------------------------------------------------------------------------------
struct CorrArgs {
  const float* a; long long lda_a;
  const float* b; long long ldb;
  int dims, mode, reduction, ncol;
  const long long* seg_start;
  const long long* seg_len;
  const long long* chunk_base;
  const int* blk_seg;
  const long long* blk_ch0;
  int chunks_per_block;
  long long g;
  const float* mu_a; const float* mu_b; const float* power; const double* lda;
  double* part;
};

// Rows of DIMS floats (DIMS % 4 == 0), 16-byte loads, two frames per lane in flight.
template <int DIMS, bool ALL_COLUMNS>
__device__ __forceinline__ void chunk_vec_rows(const CorrArgs& P, const float* A, const float* B,
                                               long long ch, const float* mu_a, const float* mu_b,
                                               const float* power, const float* rcp, const double* lda,
                                               double* out, int lane) {
  const long long g = P.g;
  double s = 0.0;
  double sd[ALL_COLUMNS ? DIMS : 1];
#pragma unroll
  for (int d = 0; d < (ALL_COLUMNS ? DIMS : 1); ++d) sd[d] = 0.0;
  const long long t_end = (ch + 1) * g;
  for (long long t = ch * g + lane; t < t_end; t += 64) {
    float av[2][DIMS], bv[2][DIMS];
    const bool second = t + 32 < t_end;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long tt = (u == 0 || second) ? t + 32 * u : t;
      const float4* pa = reinterpret_cast<const float4*>(A + tt * P.lda_a);
      const float4* pb = reinterpret_cast<const float4*>(B + tt * P.ldb);
#pragma unroll
      for (int v = 0; v < DIMS / 4; ++v) {
        const float4 x = __ldg(pa + v), y = __ldg(pb + v);
        av[u][4 * v] = x.x; av[u][4 * v + 1] = x.y; av[u][4 * v + 2] = x.z; av[u][4 * v + 3] = x.w;
        bv[u][4 * v] = y.x; bv[u][4 * v + 1] = y.y; bv[u][4 * v + 2] = y.z; bv[u][4 * v + 3] = y.w;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !second) break;
      if constexpr (!ALL_COLUMNS) {
        s += frame_signal_regs<DIMS>(av[u], bv[u], P.reduction, mu_a, mu_b, power, rcp, lda);
      } else {
        float c[DIMS];
        float mag = 0.0f;
#pragma unroll
        for (int d = 0; d < DIMS; ++d) {
          c[d] = div_fast((av[u][d] - mu_a[d]) * (bv[u][d] - mu_b[d]), power[d], rcp[d]);
          mag += fabsf(c[d]);
        }
        if (!(mag < 1e30f)) {
#pragma unroll
          for (int d = 0; d < DIMS; ++d) c[d] = ((av[u][d] - mu_a[d]) * (bv[u][d] - mu_b[d])) / power[d];
        }
#pragma unroll
        for (int d = 0; d < DIMS; ++d) sd[d] += (double)c[d];
      }
    }
  }
  if constexpr (!ALL_COLUMNS) {
    s = warp_sum(s);
    if (lane == 0) out[0] = s;
  } else {
#pragma unroll
    for (int d = 0; d < DIMS; ++d) {
      const double v = warp_sum(sd[d]);
      if (lane == 0) out[d] = v;
    }
  }
}

// One correlation column, contiguous samples (lda == ldb == 1): four frames per 16-byte load,
// two loads per lane in flight.  mode 1, reduction 'first'.
__device__ __forceinline__ void chunk_vec_scalar(const CorrArgs& P, const float* A, const float* B,
                                                 long long ch, float mu_a, float mu_b, float power,
                                                 float rcp, double* out, int lane) {
  const long long g = P.g;
  const float4* pa = reinterpret_cast<const float4*>(A + ch * g);
  const float4* pb = reinterpret_cast<const float4*>(B + ch * g);
  const long long nv = g >> 2;
  double s = 0.0;
  for (long long i = lane; i < nv; i += 64) {
    const bool second = i + 32 < nv;
    const float4 x0 = __ldg(pa + i), y0 = __ldg(pb + i);
    const float4 x1 = __ldg(pa + (second ? i + 32 : i)), y1 = __ldg(pb + (second ? i + 32 : i));
    // float32 per frame in the reference's order; the frames are carried in float64 one by one
    float c[8];
    c[0] = div_fast((x0.x - mu_a) * (y0.x - mu_b), power, rcp);
    c[1] = div_fast((x0.y - mu_a) * (y0.y - mu_b), power, rcp);
    c[2] = div_fast((x0.z - mu_a) * (y0.z - mu_b), power, rcp);
    c[3] = div_fast((x0.w - mu_a) * (y0.w - mu_b), power, rcp);
    c[4] = div_fast((x1.x - mu_a) * (y1.x - mu_b), power, rcp);
    c[5] = div_fast((x1.y - mu_a) * (y1.y - mu_b), power, rcp);
    c[6] = div_fast((x1.z - mu_a) * (y1.z - mu_b), power, rcp);
    c[7] = div_fast((x1.w - mu_a) * (y1.w - mu_b), power, rcp);
    float mag = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) mag += fabsf(c[e]);
    if (!(mag < 1e30f)) {
      c[0] = ((x0.x - mu_a) * (y0.x - mu_b)) / power; c[1] = ((x0.y - mu_a) * (y0.y - mu_b)) / power;
      c[2] = ((x0.z - mu_a) * (y0.z - mu_b)) / power; c[3] = ((x0.w - mu_a) * (y0.w - mu_b)) / power;
      c[4] = ((x1.x - mu_a) * (y1.x - mu_b)) / power; c[5] = ((x1.y - mu_a) * (y1.y - mu_b)) / power;
      c[6] = ((x1.z - mu_a) * (y1.z - mu_b)) / power; c[7] = ((x1.w - mu_a) * (y1.w - mu_b)) / power;
    }
    s += (double)c[0]; s += (double)c[1]; s += (double)c[2]; s += (double)c[3];
    if (second) { s += (double)c[4]; s += (double)c[5]; s += (double)c[6]; s += (double)c[7]; }
  }
  s = warp_sum(s);
  if (lane == 0) out[0] = s;
}

// PATH 0: generic (any dims / strides / g);  1 / 3: rows of DIMS floats, 16-byte loads (host checked
// the alignment of every segment), reduced signal / every column;  2: one contiguous column, four
// frames per load.  Separate
// instantiations so that each path gets its own register allocation (occupancy).
template <int PATH, int DIMS>
__global__ void __launch_bounds__(CORR_THREADS, PATH == 0 ? 1 : (DIMS <= 8 ? 3 : 2))
corr_partials_kernel(const CorrArgs P) {
  const int seg = P.blk_seg[blockIdx.x];
  const long long ch0 = P.blk_ch0[blockIdx.x];
  const long long n = P.seg_len[seg];
  const long long g = P.g;
  const long long nchunk = n / g;
  const long long ch1 = min(nchunk, ch0 + (long long)P.chunks_per_block);
  const int dims = P.dims, mode = P.mode, reduction = P.reduction, ncol = P.ncol;
  const float* A = P.a + P.seg_start[seg] * P.lda_a;
  const float* B = P.b ? P.b + P.seg_start[seg] * P.ldb : nullptr;
  double* part = P.part + P.chunk_base[seg] * ncol;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;

  // Global statistics (and LDA axis) staged once per block.
  __shared__ float s_mu_a[128], s_mu_b[128], s_power[128], s_rcp[128];
  __shared__ double s_lda[130];
  const float *mu_a = P.mu_a, *mu_b = P.mu_b, *power = P.power;
  const float* rcp = s_rcp;
  const double* lda = P.lda;
  if (mode == 1 || mode == 3) {
    for (int d = threadIdx.x; d < dims; d += blockDim.x) {
      const float pw = power[d];
      s_mu_a[d] = mu_a[d]; s_mu_b[d] = mu_b[d]; s_power[d] = pw;
      s_rcp[d] = (fabsf(pw) >= 1e-30f && fabsf(pw) <= 1e30f) ? 1.0f / pw : __int_as_float(0x7fc00000);
    }
    if (lda) for (int d = threadIdx.x; d < dims + 2; d += blockDim.x) s_lda[d] = lda[d];
    __syncthreads();
    mu_a = s_mu_a; mu_b = s_mu_b; power = s_power;
    if (lda) lda = s_lda;
  }

  if constexpr (PATH == 1 || PATH == 3) {
    for (long long ch = ch0 + warp; ch < ch1; ch += nwarp)
      chunk_vec_rows<DIMS, PATH == 3>(P, A, B, ch, mu_a, mu_b, power, rcp, lda, part + ch * ncol, lane);
  } else if constexpr (PATH == 2) {
    for (long long ch = ch0 + warp; ch < ch1; ch += nwarp)
      chunk_vec_scalar(P, A, B, ch, mu_a[0], mu_b[0], power[0], rcp[0], part + ch, lane);
  } else if (g >= 32) {
    for (long long ch = ch0 + warp; ch < ch1; ch += nwarp) {
      if (mode < 2) {
        double s = 0.0;
        for (long long t = ch * g + lane; t < (ch + 1) * g; t += 32)
          s += frame_signal(A + t * P.lda_a, B ? B + t * P.ldb : nullptr, dims, mode, reduction,
                            mu_a, mu_b, power, lda);
        s = warp_sum(s);
        if (lane == 0) part[ch] = s;
      } else if (mode == 3) {          // every column of the global-statistics signal
        for (int d = 0; d < dims; ++d) {
          double s = 0.0;
          for (long long t = ch * g + lane; t < (ch + 1) * g; t += 32)
            s += (double)(((A[t * P.lda_a + d] - mu_a[d]) * (B[t * P.ldb + d] - mu_b[d])) / power[d]);
          s = warp_sum(s);
          if (lane == 0) part[ch * ncol + d] = s;
        }
      } else {
        for (int d = 0; d < dims; ++d) {
          double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
          for (long long t = ch * g + lane; t < (ch + 1) * g; t += 32) {
            const double x = (double)A[t * P.lda_a + d], y = (double)B[t * P.ldb + d];
            sa += x; sb += y; saa += x * x; sbb += y * y; sab += x * y;
          }
          sa = warp_sum(sa); sb = warp_sum(sb); saa = warp_sum(saa); sbb = warp_sum(sbb);
          sab = warp_sum(sab);
          if (lane == 0) {
            double* p = part + ch * ncol + 5 * d;
            p[0] = sa; p[1] = sb; p[2] = saa; p[3] = sbb; p[4] = sab;
          }
        }
      }
    }
  } else {
    for (long long ch = ch0 + threadIdx.x; ch < ch1; ch += blockDim.x) {
      if (mode < 2) {
        double s = 0.0;
        for (long long t = ch * g; t < (ch + 1) * g; ++t)
          s += frame_signal(A + t * P.lda_a, B ? B + t * P.ldb : nullptr, dims, mode, reduction,
                            mu_a, mu_b, power, lda);
        part[ch] = s;
      } else if (mode == 3) {
        for (int d = 0; d < dims; ++d) {
          double s = 0.0;
          for (long long t = ch * g; t < (ch + 1) * g; ++t)
            s += (double)(((A[t * P.lda_a + d] - mu_a[d]) * (B[t * P.ldb + d] - mu_b[d])) / power[d]);
          part[ch * ncol + d] = s;
        }
      } else {
        for (int d = 0; d < dims; ++d) {
          double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
          for (long long t = ch * g; t < (ch + 1) * g; ++t) {
            const double x = (double)A[t * P.lda_a + d], y = (double)B[t * P.ldb + d];
            sa += x; sb += y; saa += x * x; sbb += y * y; sab += x * y;
          }
          double* p = part + ch * ncol + 5 * d;
          p[0] = sa; p[1] = sb; p[2] = saa; p[3] = sbb; p[4] = sab;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Phase 2: every window of every shape is a short sum of consecutive partials.  A block handles
// 64 windows of ONE (shape, segment) pair (table blk_j / blk_seg / blk_w0).
//   mode 0/1/3: one thread per (window, column);  mode 2: one warp per window (lanes over columns,
//   then the all-columns zero rule of brain_model.py:72-79).
// ---------------------------------------------------------------------------------------------
constexpr int WIN_PER_BLOCK = 64;

struct WinArgs {
  const double* part;
  const long long* seg_len;
  const long long* chunk_base;
  const long long* win_base;       // [n_sizes * n_seg]
  const int* blk_j;
  const int* blk_seg;
  const long long* blk_w0;
  WinSpec ws;
  int n_seg, dims, mode, ncol, out_cols;
  double* out;
};

__global__ void __launch_bounds__(CORR_THREADS)
corr_windows_kernel(const WinArgs P) {
  const int j = P.blk_j[blockIdx.x], seg = P.blk_seg[blockIdx.x];
  const long long w0 = P.blk_w0[blockIdx.x];
  const long long n = P.seg_len[seg];
  const long long W = P.ws.W[j], S = P.ws.S[j], g = P.ws.g;
  const long long nwin = (n - W) / S + 1;
  const long long w1 = min(nwin, w0 + (long long)WIN_PER_BLOCK);
  const long long cw = W / g, cs = S / g;
  const long long base = P.win_base[(long long)j * P.n_seg + seg];
  const int ncol = P.ncol, dims = P.dims, out_cols = P.out_cols;
  const double* part = P.part + P.chunk_base[seg] * ncol;
  double* out = P.out;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  if (P.mode < 2) {
    for (long long i = w0 + threadIdx.x; i < w1; i += blockDim.x) {
      double s = 0.0;
      for (long long c = i * cs; c < i * cs + cw; ++c) s += part[c];
      out[(base + i) * out_cols] = s / (double)W;
    }
  } else if (P.mode == 3) {
    for (long long e = threadIdx.x; e < (w1 - w0) * dims; e += blockDim.x) {
      const long long i = w0 + e / dims;
      const int d = (int)(e % dims);
      double s = 0.0;
      for (long long c = i * cs; c < i * cs + cw; ++c) s += part[c * ncol + d];
      out[(base + i) * out_cols + d] = s / (double)W;
    }
  } else {
    for (long long i = w0 + warp; i < w1; i += nwarp) {
      double prod_x = 1.0, prod_y = 1.0;
      for (int d0 = 0; d0 < dims; d0 += 32) {
        const int d = d0 + lane;
        double r = 0.0, xp = 1.0, yp = 1.0;
        if (d < dims) {
          double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
          for (long long c = i * cs; c < i * cs + cw; ++c) {
            const double* p = part + c * ncol + 5 * d;
            sa += p[0]; sb += p[1]; saa += p[2]; sbb += p[3]; sab += p[4];
          }
          const double inv = 1.0 / (double)W;
          xp = saa - sa * sa * inv;
          yp = sbb - sb * sb * inv;
          // A constant column has zero power in the reference's centred float32 sums; raw
          // float64 moments leave rounding dust instead.
          if (xp <= saa * 1e-13) xp = 0.0;
          if (yp <= sbb * 1e-13) yp = 0.0;
          r = (sab - sa * sb * inv) / (sqrt(xp) * sqrt(yp));
          out[(base + i) * out_cols + d] = r;
        }
        // product over columns of the powers (brain_model.py:72-73)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          xp *= __shfl_xor_sync(0xffffffffu, xp, o);
          yp *= __shfl_xor_sync(0xffffffffu, yp, o);
        }
        prod_x *= xp;
        prod_y *= yp;
      }
      if (!(prod_x > 0.0) || !(prod_y > 0.0)) {
        __syncwarp();
        for (int d = lane; d < dims; d += 32) out[(base + i) * out_cols + d] = 0.0;
      }
    }
  }
}

__global__ void column_moments_kernel(const float* __restrict__ a, long long lda,
                                      const float* __restrict__ b, long long ldb, long long n,
                                      int dims, double* sums) {
  // threadIdx.x -> column (strided), threadIdx.y + blockIdx -> rows
  __shared__ double red[4][8][33];
  for (int d0 = 0; d0 < dims; d0 += 32) {
    const int d = d0 + threadIdx.x;
    double sa = 0, sb = 0, saa = 0, sbb = 0;
    if (d < dims) {
      for (long long t = (long long)blockIdx.x * blockDim.y + threadIdx.y; t < n;
           t += (long long)gridDim.x * blockDim.y) {
        const double x = (double)a[t * lda + d], y = (double)b[t * ldb + d];
        sa += x; sb += y; saa += x * x; sbb += y * y;
      }
    }
    red[0][threadIdx.y][threadIdx.x] = sa;
    red[1][threadIdx.y][threadIdx.x] = sb;
    red[2][threadIdx.y][threadIdx.x] = saa;
    red[3][threadIdx.y][threadIdx.x] = sbb;
    __syncthreads();
    if (threadIdx.y < 4 && d < dims) {
      double s = 0;
      for (int y = 0; y < 8; ++y) s += red[threadIdx.y][y][threadIdx.x];
      atomicAdd(sums + (long long)threadIdx.y * dims + d, s);
    }
    __syncthreads();
  }
}

__global__ void wta_kernel(const double* __restrict__ r1, const double* __restrict__ r2, long long n,
                           uint8_t* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (r1[i] > r2[i]) ? 1 : 0;
}

}  // namespace
}  // namespace td

using namespace td;

extern "C" int64_t td_window_count(int64_t n_frames, int64_t W, int64_t S) {
  if (W <= 0 || S <= 0 || n_frames < W) return 0;
  return (n_frames - W) / S + 1;
}

extern "C" int td_column_moments_f64(const float* a, int64_t lda, const float* b, int64_t ldb,
                                     int64_t n, int dims, double* sums, void* stream) {
  TD_REQUIRE(a && b && sums && dims > 0 && n >= 0 && lda >= dims && ldb >= dims,
             "td_column_moments_f64: bad arguments");
  if (n == 0) return TD_OK;
  const int blocks = (int)std::min<int64_t>(1184, ceil_div(n, 8));
  column_moments_kernel<<<blocks, dim3(32, 8), 0, (cudaStream_t)stream>>>(a, lda, b, ldb, n, dims, sums);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}

extern "C" int td_windowed_corr_f32(const float* a, int64_t lda, const float* b, int64_t ldb, int dims,
                                    const int64_t* seg_start, const int64_t* seg_len, int n_seg,
                                    const int64_t* W, const int64_t* S, int n_sizes, int mode,
                                    int reduction, const float* mu_a, const float* mu_b,
                                    const float* power, const double* lda_params, double* out,
                                    int64_t* out_offset, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(a && seg_start && seg_len && W && S && n_seg > 0, "td_windowed_corr_f32: null argument");
  TD_REQUIRE(n_sizes >= 1 && n_sizes <= MAX_SIZES, "td_windowed_corr_f32: n_sizes must be in [1, %d]", MAX_SIZES);
  TD_REQUIRE(mode >= 0 && mode <= 3, "td_windowed_corr_f32: unknown mode %d", mode);
  TD_REQUIRE(dims >= 1 && lda >= (mode == 0 ? 1 : dims), "td_windowed_corr_f32: bad dims/lda");
  if (mode != 0) TD_REQUIRE(b && ldb >= dims, "td_windowed_corr_f32: second signal missing");
  if (mode == 1 || mode == 3) {
    TD_REQUIRE(mu_a && mu_b && power, "td_windowed_corr_f32: global statistics missing");
    TD_REQUIRE((reduction >= 0 && reduction <= 3) || reduction == 5, "td_windowed_corr_f32: unknown reduction %d", reduction);
    TD_REQUIRE(reduction != 5 || lda_params, "td_windowed_corr_f32: LDA parameters missing");
    TD_REQUIRE(reduction != 1 || dims >= 2, "td_windowed_corr_f32: 'second' needs >= 2 columns");
    TD_REQUIRE(dims <= 128, "td_windowed_corr_f32: at most 128 correlation columns");
  }
  WinSpec ws;
  ws.n_sizes = n_sizes;
  long long g = 0;
  for (int j = 0; j < n_sizes; ++j) {
    TD_REQUIRE(W[j] >= 1 && S[j] >= 1, "td_windowed_corr_f32: window %d has width %lld step %lld (step 0 never terminates in the reference either)",
               j, (long long)W[j], (long long)S[j]);
    ws.W[j] = W[j];
    ws.S[j] = S[j];
    g = std::gcd(g, (long long)W[j]);
    g = std::gcd(g, (long long)S[j]);
  }
  ws.g = g;
  const int ncol = (mode == 2) ? 5 * dims : (mode == 3 ? dims : 1);
  const int out_cols = (mode >= 2) ? dims : 1;
  std::vector<long long> win_base((size_t)n_sizes * n_seg);
  int64_t total = 0;
  for (int j = 0; j < n_sizes; ++j) {
    if (out_offset) out_offset[j] = total;
    for (int s = 0; s < n_seg; ++s) {
      win_base[(size_t)j * n_seg + s] = total;
      total += td_window_count(seg_len[s], W[j], S[j]);
    }
  }
  if (out_offset) out_offset[n_sizes] = total;
  if (total == 0) return TD_OK;
  TD_REQUIRE(out != nullptr, "td_windowed_corr_f32: null output");

  // Phase-1 blocks: consecutive chunks of one segment.  One warp per chunk when g >= 32 (a few
  // chunks per warp keep the grid near 16 blocks per SM), one thread per chunk otherwise.
  std::vector<long long> chunk_base(n_seg);
  long long n_chunks = 0;
  for (int s = 0; s < n_seg; ++s) { chunk_base[s] = n_chunks; n_chunks += seg_len[s] / g; }
  int cpb;
  if (g >= 32) {
    const long long want_blocks = 148LL * 16;
    cpb = (int)std::min<long long>(64, std::max<long long>(CORR_THREADS / 32, ceil_div(n_chunks, want_blocks)));
    cpb = (int)round_up(cpb, CORR_THREADS / 32);
  } else {
    cpb = CORR_THREADS * 4;
  }
  std::vector<int> blk_seg, wblk_j, wblk_seg;
  std::vector<long long> blk_ch0, wblk_w0;
  for (int s = 0; s < n_seg; ++s)
    for (long long c = 0; c < seg_len[s] / g; c += cpb) { blk_seg.push_back(s); blk_ch0.push_back(c); }
  for (int j = 0; j < n_sizes; ++j)
    for (int s = 0; s < n_seg; ++s) {
      const long long nw = td_window_count(seg_len[s], W[j], S[j]);
      for (long long w0 = 0; w0 < nw; w0 += WIN_PER_BLOCK) { wblk_j.push_back(j); wblk_seg.push_back(s); wblk_w0.push_back(w0); }
    }
  DevBufs bufs(st);
  std::vector<long long> ss(seg_start, seg_start + n_seg), sl(seg_len, seg_len + n_seg);
  long long *ss_d, *sl_d, *wb_d, *cb_d, *bc_d, *ww_d;
  int *bs_d, *wj_d, *wsg_d;
  double* part_d;
  int rc;
  if ((rc = bufs.put(ss, &ss_d)) || (rc = bufs.put(sl, &sl_d)) || (rc = bufs.put(win_base, &wb_d)) ||
      (rc = bufs.put(chunk_base, &cb_d)) || (rc = bufs.put(blk_seg, &bs_d)) || (rc = bufs.put(blk_ch0, &bc_d)) ||
      (rc = bufs.put(wblk_j, &wj_d)) || (rc = bufs.put(wblk_seg, &wsg_d)) || (rc = bufs.put(wblk_w0, &ww_d)) ||
      (rc = bufs.alloc((size_t)std::max<long long>(1, n_chunks) * ncol, &part_d)))
    return rc;
  CorrArgs ca;
  ca.a = a; ca.lda_a = lda; ca.b = b; ca.ldb = ldb;
  ca.dims = dims; ca.mode = mode; ca.reduction = reduction; ca.ncol = ncol;
  ca.seg_start = ss_d; ca.seg_len = sl_d; ca.chunk_base = cb_d; ca.blk_seg = bs_d; ca.blk_ch0 = bc_d;
  ca.chunks_per_block = cpb; ca.g = g;
  ca.mu_a = mu_a; ca.mu_b = mu_b; ca.power = power; ca.lda = lda_params; ca.part = part_d;
  if (!blk_seg.empty()) {
    // Vector paths need 16-byte aligned rows in every segment.
    bool aligned = b != nullptr;
    for (int s = 0; s < n_seg && aligned; ++s)
      aligned = ((reinterpret_cast<uintptr_t>(a + seg_start[s] * lda) |
                  reinterpret_cast<uintptr_t>(b + seg_start[s] * ldb)) % 16) == 0;
    const bool rows_vec = aligned && (mode == 1 || mode == 3) && g >= 32 && (dims == 4 || dims == 8 || dims == 16) &&
                          lda % 4 == 0 && ldb % 4 == 0;
    const bool scalar_vec = aligned && mode == 1 && reduction == 0 && dims == 1 && g >= 128 && g % 4 == 0 &&
                            lda == 1 && ldb == 1;
    const unsigned nblk = (unsigned)blk_seg.size();
    if (rows_vec && mode == 1 && dims == 4) corr_partials_kernel<1, 4><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else if (rows_vec && mode == 1 && dims == 8) corr_partials_kernel<1, 8><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else if (rows_vec && mode == 1) corr_partials_kernel<1, 16><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else if (rows_vec && dims == 4) corr_partials_kernel<3, 4><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else if (rows_vec && dims == 8) corr_partials_kernel<3, 8><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else if (rows_vec) corr_partials_kernel<3, 16><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else if (scalar_vec) corr_partials_kernel<2, 1><<<nblk, CORR_THREADS, 0, st>>>(ca);
    else corr_partials_kernel<0, 1><<<nblk, CORR_THREADS, 0, st>>>(ca);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  WinArgs wa;
  wa.part = part_d; wa.seg_len = sl_d; wa.chunk_base = cb_d; wa.win_base = wb_d;
  wa.blk_j = wj_d; wa.blk_seg = wsg_d; wa.blk_w0 = ww_d; wa.ws = ws;
  wa.n_seg = n_seg; wa.dims = dims; wa.mode = mode; wa.ncol = ncol; wa.out_cols = out_cols; wa.out = out;
  corr_windows_kernel<<<(unsigned)wblk_j.size(), CORR_THREADS, 0, st>>>(wa);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}

extern "C" int td_wta_decide_u8(const double* r1, const double* r2, int64_t n, uint8_t* out, void* stream) {
  TD_REQUIRE(r1 && r2 && out && n >= 0, "td_wta_decide_u8: bad arguments");
  if (n == 0) return TD_OK;
  wta_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(r1, r2, n, out);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}
