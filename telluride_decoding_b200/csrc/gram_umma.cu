// Lagged (implicit-Hankel) Gram on tcgen05 tensor cores, split operands (fp16 x 3 / 3xTF32), for sm_100a.
//
// The default for designs of >= 512 columns is hk::lagged_gram_halo_kernel (gram_halo.cuh, methods 6
// and 4: the lagged operand is built on chip from a raw TMA tile + lag halo); its host side -- plan,
// tables, the pre-split copy of the fp16 x 3 flavour, the block-Toeplitz method 5 -- is in this file
// below the round-1 kernels, which are described first and remain for narrow designs (method 2 with
// single CTAs) and as selectable flavours (methods 2 and 3):
//
// Replaces the reference's minibatch loop `sum_xtx += np.matmul(x.T, x)` over the materialised
// lagged design (telluride_decoding/brain_model.py:429-444, cca.py:311-332; the design itself is
// built by brain_data.py:446-455).  Here the design Z = [X_0 | X_1] never exists:
//
//   1. stage:  every raw stream [T, C] (time-major) is rewritten once channel-major (4 copies,
//      advanced by 0..3 frames, so that every lag is a 16-byte aligned TMA box origin), segments
//      separated by >= (pre+post) zero frames.  Row (stream q, lag k, channel c) of Z' is then the
//      staged row c read at a time shift of (k - pre): a plain TMA box at a shifted coordinate.
//      Zero padding at segment edges comes from the zero gaps / TMA out-of-bounds fill.
//   2. gram:   persistent warp-specialised kernels.  K dimension = time, cut into 32-frame k-tiles;
//      M/N = rows of Z' in 128-row operand tiles made of CB-channel TMA boxes; raw fp32 tiles land
//      by TMA (K-major, 128B swizzle); the 3xTF32 split is built on chip (hi = the raw tile, because
//      kind::tf32 truncates the 13 low mantissa bits -- verified on B200; lo = tf32_rna(x - trunc(x)));
//      hi*hi + hi*lo + lo*hi tcgen05.mma into TMEM; every `chunk` k-tiles the accumulator is drained
//      into fp32 registers (bounds the tensor-core accumulation length).  Two flavours:
//        * ts::lagged_gram_ts_kernel (gram_ts.cuh): CTA pairs, A operand written to TENSOR memory by
//          the converter warps, two MMA issuing warps, three rotating accumulator regions (method 3).
//        * lagged_gram_umma_kernel<CG> (below): both operands from shared memory.
//            warp 0      TMA producer, 3-stage ring
//            warps 10-11 converters (lo tile + masking of frames that are not design rows)
//            warp 1      MMA issue (elect.sync), double-buffered TMEM accumulator
//            warps 2-9   epilogue
//          CG = 1 (128x128 tiles) for narrow designs; CG = 2 kept as method 2.
//      Work = (superchunk, lower-triangle tile, k-range) split evenly over CTAs (stream-K); partial
//      tiles are combined with fp32 red.add into G.
//   3. the computed triangle is mirrored by td_symmetrize_f32.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace td {

int launch_gram_simt(const td_stream* streams, int n_streams, const int64_t* seg_first,
                     const int64_t* seg_keep, int n_seg, float* G, int64_t ldg, cudaStream_t st);

namespace {

constexpr int BK = 32;                       // frames per k-tile (one 128-byte swizzle row)
constexpr int TILE_ROWS = 128;               // operand rows per CTA
constexpr int STAGES = 3;
constexpr int OPERAND_BYTES = TILE_ROWS * BK * 4;       // 16 KiB
constexpr int STAGE_BYTES = 4 * OPERAND_BYTES;          // A_raw(hi), A_lo, B_raw(hi), B_lo
constexpr int TMA_BYTES = 2 * OPERAND_BYTES;            // what the producer lands per stage
#ifndef TD_CONV_WARPS
#define TD_CONV_WARPS 2
#endif
constexpr int NUM_CONV_WARPS = TD_CONV_WARPS;     // converter warps (half of them per operand)
constexpr int NUM_THREADS = 320 + 32 * NUM_CONV_WARPS;   // warp 0 TMA, warp 1 MMA/TMEM, warps 2..9 epilogue, 10.. converters
constexpr int NUM_EPI_WARPS = 8;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(8 * (4 * STAGES + 4) + 8 <= 256, "barrier block too small");

struct GramParams {
  const int4* groups;     // [n_groups] {stream, aligned time shift, first channel, shift plane}
  const int2* ktiles;     // [NK] {staged time of design row 0 of the tile, begin | end << 8}: rows [begin, end) exist
  const int* out_idx;     // [Dpad] tile-space row -> row/col of G, -1 for padding
  float* G;
  long long ldg;
  int CB;                 // channels per TMA box
  int groups_per_tile;    // TILE_ROWS / CB
  int ntri;               // lower-triangle tile count
  long long NK;           // k-tiles in total
  int KS;                 // k-tiles per superchunk
  int chunk_kt;           // k-tiles per TMEM flush
  long long total_steps;  // NSC * ntri * KS
  long long* prof;        // development: per-CTA stall clocks of the TS kernel (null in production)
};

struct WorkItem {
  int tm, tn;
  long long kt_begin, kt_end;
};

// Work = (superchunk, lower-triangle tile, k-tile).  All workers sweep the SAME superchunk at the
// same time (its staged frames stay L2 resident while every tile consumes them): inside superchunk
// sc the ntri * KS k-tile steps are split evenly over the workers (stream-K), then everyone moves
// on to sc + 1.  Identical decode in every warp role.
// Number of A-operand tiles of a RECTANGULAR problem (A rows x all B rows; method 5), 0 for the
// symmetric lower-triangle tile set.
template <class PT>
__device__ __forceinline__ int rect_tiles_of(const PT&) { return 0; }

struct WorkIter {
  long long sc, nsc, per_sc, p, hi;
  int worker, n_workers;
  template <class PT>
  __device__ WorkIter(const PT& P, int worker_, int n_workers_) {
    worker = worker_;
    n_workers = n_workers_;
    per_sc = (long long)P.ntri * P.KS;
    nsc = P.total_steps / per_sc;
    sc = -1;
    p = hi = 0;
  }
  template <class PT>
  __device__ bool next(const PT& P, WorkItem& it) {
    for (;;) {
      if (p >= hi) {
        if (++sc >= nsc) return false;
        const int w = (int)((worker + sc) % n_workers);     // rotate who gets the ragged pieces
        p = per_sc * w / n_workers;
        hi = per_sc * (w + 1) / n_workers;
        continue;
      }
      const int tile = (int)(p / P.KS);
      const int kt0 = (int)(p - (long long)tile * P.KS);
      const long long len = min((long long)(P.KS - kt0), hi - p);
      const long long b = sc * P.KS + kt0;
      const long long e = min(b + len, P.NK);
      p += len;
      if (b < e) {
        const int ta = rect_tiles_of(P);
        if (ta > 0) {                              // rectangular: tile = tn * ta + tm
          it.tn = tile / ta;
          it.tm = tile - it.tn * ta;
        } else {
          int tm = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
          while ((long long)(tm + 1) * (tm + 2) / 2 <= tile) ++tm;
          while ((long long)tm * (tm + 1) / 2 > tile) --tm;
          it.tm = tm;
          it.tn = tile - tm * (tm + 1) / 2;
        }
        it.kt_begin = b;
        it.kt_end = e;
        return true;
      }
    }
  }
};

// One converter warp turns the raw tile of one operand (16 KiB, K-major, 128-byte swizzle: frame t
// of row r sits in 16-byte chunk (t/4) ^ (r%8)) into its lo tile, lo = tf32_rna(x - trunc_tf32(x)),
// at the same offsets.  The raw tile doubles as the hi operand: kind::tf32 reads only the upper 19
// bits.  Frames outside [begin, end) are not design rows of this call (end of trial, rows lost to
// drop_remainder, brain_data.py:369-370, or another rank's share): they are zeroed in both tiles.
__device__ __forceinline__ void split_tile(uint32_t raw, uint32_t lo, int begin, int end, int lane,
                                           int part, int n_parts) {
  const bool partial = begin > 0 || end < BK;
  const int per_part = OPERAND_BYTES / 16 / 32 / n_parts;
#pragma unroll 4
  for (int i = part * per_part; i < (part + 1) * per_part; ++i) {
    const uint32_t off = (uint32_t)(i * 32 + lane) * 16u;
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(raw + off));
    if (partial) {
      const uint32_t r = off >> 7;
      const int t0 = (int)((((off >> 4) & 7u) ^ (r & 7u)) << 2);
      const bool k0 = t0 >= begin && t0 < end, k1 = t0 + 1 >= begin && t0 + 1 < end;
      const bool k2 = t0 + 2 >= begin && t0 + 2 < end, k3 = t0 + 3 >= begin && t0 + 3 < end;
      if (!(k0 && k1 && k2 && k3)) {
        v.x = k0 ? v.x : 0.0f; v.y = k1 ? v.y : 0.0f; v.z = k2 ? v.z : 0.0f; v.w = k3 ? v.w : 0.0f;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raw + off), "f"(v.x), "f"(v.y),
                     "f"(v.z), "f"(v.w) : "memory");
      }
    }
    float4 l;
    {
      // x - trunc(x) is exact in fp32; adding half an ulp of tf32 to the magnitude bits and
      // clearing the 13 low bits is round-to-nearest (ties away), i.e. cvt.rna.tf32.f32.
      const float dx = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
      const float dy = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
      const float dz = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
      const float dw = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
      l.x = __uint_as_float((__float_as_uint(dx) + 0x1000u) & 0xFFFFE000u);
      l.y = __uint_as_float((__float_as_uint(dy) + 0x1000u) & 0xFFFFE000u);
      l.z = __uint_as_float((__float_as_uint(dz) + 0x1000u) & 0xFFFFE000u);
      l.w = __uint_as_float((__float_as_uint(dw) + 0x1000u) & 0xFFFFE000u);
    }
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + off), "f"(l.x), "f"(l.y),
                 "f"(l.z), "f"(l.w) : "memory");
  }
  ptx::fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor-core proxy
  __syncwarp();
}

#include "gram_ts.cuh"
#include "gram_halo.cuh"

// CG = 1: one CTA per 128x128 tile.  CG = 2: a CTA pair (cluster of 2) per 256x256 tile, each CTA
// staging 128 rows of A and 128 rows of B; the pair leader issues tcgen05.mma.cta_group::2.
template <int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lagged_gram_umma_kernel(const __grid_constant__ CUtensorMap tmap0,
                        const __grid_constant__ CUtensorMap tmap1, const GramParams P) {
  constexpr int BN = TILE_ROWS * CG;          // accumulator columns (N of the MMA)
  constexpr int NACC = BN / 2;                // accumulator values per epilogue thread
  constexpr uint32_t TMEM_COLS = 2 * BN;      // double-buffered accumulator
  constexpr uint32_t IDESC = ptx::umma_idesc_tf32(128 * CG, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                 // raw tiles landed (TMA)
  auto conv_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };      // lo tiles built
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); }; // MMAs of the stage retired
  auto pready_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + s); };  // leader: peer stage ready
  auto tfull_bar = [&](int b) { return bar_base + 8u * (4 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (4 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (4 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const int worker = blockIdx.x / CG;
  const int n_workers = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap0);
    ptx::prefetch_tmap(&tmap1);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(conv_bar(s), NUM_CONV_WARPS);
      ptx::mbar_init(empty_bar(s), 1);
      ptx::mbar_init(pready_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), NUM_EPI_WARPS * CG);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<CG>(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) ==============
    // (Issuing under `if (lane == 0)` makes ptxas wrap every TMA / MMA in an elect + R2UR.BROADCAST
    // loop, ~80 clocks per instruction; under elect.sync the operands go straight to uniform registers.)
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    int stage = 0;
    uint32_t phase = 0;
    while (wi.next(P, it)) {
      const int ga = (it.tm * CG + (int)rank) * P.groups_per_tile;
      const int gb = (it.tn * CG + (int)rank) * P.groups_per_tile;
      int kt_time = __ldg(&P.ktiles[it.kt_begin]).x;
      for (long long kt = it.kt_begin; kt < it.kt_end; ++kt) {
        const int t0 = kt_time;
        if (kt + 1 < it.kt_end) kt_time = __ldg(&P.ktiles[kt + 1]).x;
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t fb = full_bar(stage);
          ptx::mbar_expect_tx(fb, TMA_BYTES);
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
#pragma unroll 1
          for (int op = 0; op < 2; ++op) {
            const int g0 = op ? gb : ga;
            const uint32_t obase = sbase + op * 2 * OPERAND_BYTES;
#pragma unroll 1
            for (int g = 0; g < P.groups_per_tile; ++g) {
              const int4 gd = __ldg(&P.groups[g0 + g]);
              ptx::tma_load_3d(obase + g * P.CB * (BK * 4), gd.x ? &tmap1 : &tmap0, fb, t0 + gd.y,
                               gd.z, gd.w);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 2 + NUM_EPI_WARPS) {
    // ===================== converters: raw tile -> lo tile (+ row masking) =====================
    constexpr int PER_OP = NUM_CONV_WARPS / 2;
    const int cw = warp - (2 + NUM_EPI_WARPS);
    const int op = cw / PER_OP;                         // 0: A operand, 1: B operand
    const int part = cw % PER_OP;
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    int stage = 0;
    uint32_t phase = 0;
    while (wi.next(P, it)) {
      for (long long kt = it.kt_begin; kt < it.kt_end; ++kt) {
        const int range = __ldg(&P.ktiles[kt]).y;
        const uint32_t raw = smem_base + stage * STAGE_BYTES + op * 2 * OPERAND_BYTES;
        ptx::mbar_wait(full_bar(stage), phase);
        split_tile(raw, raw + OPERAND_BYTES, range & 0xff, range >> 8, lane, part, PER_OP);
        if (lane == 0) ptx::mbar_arrive(conv_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA warp =====================
    // With CG == 2 the peer reports each converted stage to the leader (remote mbarrier arrive) and
    // only the leader issues MMAs.
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t tphase[2] = {0u, 0u};
    const bool issuer = (CG == 1 || rank == 0);
    const uint32_t pready_leader0 = (CG == 2) ? ptx::mapa(pready_bar(0), 0) : 0u;
    while (wi.next(P, it)) {
      int next_range = __ldg(&P.ktiles[it.kt_begin]).y;
      for (long long c0 = it.kt_begin; c0 < it.kt_end; c0 += P.chunk_kt) {
        const long long c1 = min(it.kt_end, c0 + (long long)P.chunk_kt);
        if (issuer) {
          ptx::mbar_wait(tempty_bar(buf), tphase[buf] ^ 1u);
          ptx::tc_fence_after();
        }
        for (long long kt = c0; kt < c1; ++kt) {
          const int range = next_range;
          if (kt + 1 < it.kt_end) next_range = __ldg(&P.ktiles[kt + 1]).y;
          const int begin = range & 0xff, end = range >> 8;
          const uint32_t sbase = smem_base + stage * STAGE_BYTES;
          ptx::mbar_wait(conv_bar(stage), phase);
          if constexpr (CG == 2) {
            if (!issuer) {
              if (lane == 0) ptx::mbar_arrive_remote(pready_leader0 + 8u * stage);
            } else {
              ptx::mbar_wait(pready_bar(stage), phase);
            }
          }
          if (issuer) {
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const int ks0 = begin >> 3, ks1 = (end + 7) >> 3;
              const uint32_t d = tmem_base + (uint32_t)(buf * BN);
              const uint64_t a_hi = ptx::umma_desc_k_sw128(sbase);
              const uint64_t a_lo = ptx::umma_desc_k_sw128(sbase + OPERAND_BYTES);
              const uint64_t b_hi = ptx::umma_desc_k_sw128(sbase + 2 * OPERAND_BYTES);
              const uint64_t b_lo = ptx::umma_desc_k_sw128(sbase + 3 * OPERAND_BYTES);
              for (int ks = ks0; ks < ks1; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 2);   // 8 frames = 32 B = 2 x 16 B
                const uint32_t acc = (kt > c0 || ks > ks0) ? 1u : 0u;
                ptx::mma_tf32<CG>(d, a_hi + adv, b_hi + adv, IDESC, acc);
                ptx::mma_tf32<CG>(d, a_hi + adv, b_lo + adv, IDESC, 1u);
                ptx::mma_tf32<CG>(d, a_lo + adv, b_hi + adv, IDESC, 1u);
              }
              ptx::mma_commit<CG>(empty_bar(stage));          // stage reusable once these retire
              if (kt == c1 - 1) ptx::mma_commit<CG>(tfull_bar(buf));
            }
            __syncwarp();
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tphase[buf] ^= 1u;
        buf ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps: TMEM -> fp32 registers -> red.add =====================
    const int ew = warp - 2;
    const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad+32)
    const int half = ew >> 2;                  // column half of the accumulator
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    int buf = 0;
    uint32_t fphase[2] = {0u, 0u};
    const uint32_t tempty_leader0 = (CG == 2) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
    float acc[NACC];
    while (wi.next(P, it)) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
      for (long long c0 = it.kt_begin; c0 < it.kt_end; c0 += P.chunk_kt) {
        ptx::mbar_wait(tfull_bar(buf), fphase[buf]);
        fphase[buf] ^= 1u;
        ptx::tc_fence_after();
        const uint32_t taddr =
            tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + half * NACC);
#pragma unroll
        for (int j = 0; j < NACC / 32; ++j) {
          float v[32];
          ptx::tmem_ld32(taddr + j * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) ptx::mbar_arrive_remote(tempty_leader0 + 8u * buf);
          else ptx::mbar_arrive(tempty_bar(buf));
        }
        buf ^= 1;
      }
      // Tile (tm, tn), tm >= tn: element (m, n) is stored at G[n][m] (upper triangle) so that the
      // 32 lanes of a warp (consecutive m) write consecutive addresses.
      const int m_row = (it.tm * CG + (int)rank) * TILE_ROWS + quad * 32 + lane;
      const int m_out = __ldg(&P.out_idx[m_row]);
      const int n_row0 = it.tn * BN + half * NACC;
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        const int n_row = n_row0 + i;
        const int n_out = __ldg(&P.out_idx[n_row]);
        if (m_out >= 0 && n_out >= 0 && n_row <= m_row)
          atomicAdd(P.G + (long long)n_out * P.ldg + m_out, acc[i]);
      }
    }
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// Stage: raw [T, C] time-major  ->  planes[n_shift][Cpad][Tp] channel-major (raw fp32).
// ---------------------------------------------------------------------------------------------
// Plane s holds the stream advanced by s frames, so that a lag shift d is read from plane (d mod 4)
// at the 16-byte aligned coordinate d - (d mod 4): TMA box origins must be 16-byte aligned along
// the contiguous dimension (measured: an unaligned origin raises "illegal instruction").
__global__ void stage_kernel(StreamView s, const int64_t* __restrict__ seg_off, int seg0,
                             float* __restrict__ planes, int64_t Tp, int Cpad, int n_shift) {
  __shared__ float tile[32][33];
  const int seg = seg0 + blockIdx.y;
  const int64_t len = s.seg_len[seg];
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  if (t0 >= len) return;
  const int c0 = blockIdx.z * 32;
  const float* src = s.data + (s.seg_start[seg] + t0) * s.ld;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + threadIdx.x;
    float v = 0.0f;
    if (t0 + r < len && c < s.channels) v = src[(int64_t)r * s.ld + c];
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  const int64_t plane = (int64_t)Cpad * Tp;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const int64_t t = t0 + threadIdx.x;
    if (c < s.channels && t < len) {
      const float v = tile[threadIdx.x][r];
      const int64_t o = (int64_t)c * Tp + seg_off[seg] + t;
      for (int sh = 0; sh < n_shift; ++sh) planes[sh * plane + o - sh] = v;
    }
  }
}

__global__ void symmetrize_kernel(float* __restrict__ G, int64_t n, int64_t ldg, int from_upper,
                                  int64_t batch_stride) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;                        // handle block pairs (bi <= bj) once
  G += (int64_t)blockIdx.z * batch_stride;
  // Source block: the computed triangle's (bi, bj) if from_upper, else (bj, bi).
  const int sr = from_upper ? bi : bj, sc = from_upper ? bj : bi;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = (int64_t)sr * 32 + r, j = (int64_t)sc * 32 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && j < n) ? G[i * ldg + j] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = (int64_t)sc * 32 + r, j = (int64_t)sr * 32 + threadIdx.x;   // mirrored block
    if (i < n && j < n) {
      const bool dst_is_uncomputed = from_upper ? (i > j) : (i < j);
      if (dst_is_uncomputed) G[i * ldg + j] = tile[threadIdx.x][r];
    }
  }
}

// The halo kernel stores every unordered pair of design columns exactly once, at G[n][m] where (m, n)
// is its position in TILE space -- which may be either triangle of G, because the tile-space row
// order (stream, channel group, lag group) is not the column order of the design (lag-major).  G
// starts from zero, so G[i][j] + G[j][i] is the pair's value; it is written to both positions
// (exactly symmetric, like numpy's syrk path).
__global__ void fold_symmetric_kernel(float* __restrict__ G, int64_t n, int64_t ldg, int64_t batch_stride) {
  __shared__ float tu[32][33], tl[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  G += (int64_t)blockIdx.z * batch_stride;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t iu = (int64_t)bi * 32 + r, ju = (int64_t)bj * 32 + threadIdx.x;
    const int64_t il = (int64_t)bj * 32 + r, jl = (int64_t)bi * 32 + threadIdx.x;
    tu[r][threadIdx.x] = (iu < n && ju < n) ? G[iu * ldg + ju] : 0.0f;
    tl[r][threadIdx.x] = (il < n && jl < n) ? G[il * ldg + jl] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = threadIdx.x;
    const int64_t iu = (int64_t)bi * 32 + r, ju = (int64_t)bj * 32 + c;
    const int64_t il = (int64_t)bj * 32 + r, jl = (int64_t)bi * 32 + c;
    if (bi == bj) {
      if (iu < n && ju < n) G[iu * ldg + ju] = (r == c) ? tu[r][c] : tu[r][c] + tu[c][r];
    } else {
      if (iu < n && ju < n) G[iu * ldg + ju] = tu[r][c] + tl[c][r];
      if (il < n && jl < n) G[il * ldg + jl] = tl[r][c] + tu[c][r];
    }
  }
}

// Padded copy of a raw stream whose rows TMA cannot address (row stride or base not a multiple of
// 16 bytes): rows [row0, row0 + n_rows) -> dst[n_rows][ldp], channels beyond C are zero.
__global__ void pad_rows_kernel(const float* __restrict__ src, int64_t ld, int C, int64_t row0,
                                int64_t n_rows, float* __restrict__ dst, int ldp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * ldp) return;
  const int64_t r = i / ldp;
  const int c = (int)(i - r * ldp);
  dst[i] = (c < C) ? src[(row0 + r) * ld + c] : 0.0f;
}

// ---- fp16 x 3 flavour of the halo kernel: per-channel scale and the pre-split copy ----------------
// Largest |sample| per channel over rows [row0, row0 + n_rows) (bit pattern of a non-negative float:
// unsigned order = numeric order; a NaN ends up on top and switches the scaling off for that channel).
// block (32 channels, 8 row lanes); 1-D grid of (row slab, channel block) with the channel block fastest,
// so that the blocks sharing a slab's sectors run back to back (rows are rarely 128-byte aligned).
constexpr int SPLIT_ROWS_PER_BLOCK = 256;
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ src, int64_t ld, int C, int64_t row0,
                                                     int64_t n_rows, unsigned* __restrict__ amax_bits, int n_cblocks) {
  __shared__ unsigned red[8][32];
  const int c = (int)(blockIdx.x % n_cblocks) * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)(blockIdx.x / n_cblocks) * SPLIT_ROWS_PER_BLOCK;
  const int64_t r1 = min(r0 + SPLIT_ROWS_PER_BLOCK, n_rows);
  unsigned m = 0;
  if (c < C)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8)
      m = max(m, __float_as_uint(src[(row0 + r) * ld + c]) & 0x7fffffffu);
  red[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int i = 1; i < 8; ++i) m = max(m, red[i][threadIdx.x]);
    if (m) atomicMax(&amax_bits[c], m);
  }
}

// Power of two that moves a channel's largest sample into [2^14, 2^15) (fp16 tops out at 65504; the lo
// halves then sit 11 bits lower, far above the fp16 subnormal spacing of 2^-24).
__device__ __forceinline__ int scale_exponent(unsigned amax_bits) {
  const float a = __uint_as_float(amax_bits);
  if (!(a > 0.0f) || !(a < __int_as_float(0x7f800000))) return 0;
  int e;
  frexpf(a, &e);                                   // a = m * 2^e, m in [0.5, 1)
  return max(-120, min(120, 15 - e));
}

// dst[r][c] = fp16(x') | fp16(x' - fp16(x')) << 16 with x' = src[row0 + r][c] * 2^scale_exponent(c);
// channels in [C, ldp) are zero.  Same thread layout as absmax_kernel.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ src, int64_t ld, int C, int64_t row0,
                                                         int64_t n_rows, const unsigned* __restrict__ amax_bits,
                                                         uint32_t* __restrict__ dst, int ldp, int n_cblocks) {
  const int c = (int)(blockIdx.x % n_cblocks) * 32 + threadIdx.x;
  if (c >= ldp) return;
  const int64_t r0 = (int64_t)(blockIdx.x / n_cblocks) * SPLIT_ROWS_PER_BLOCK;
  const int64_t r1 = min(r0 + SPLIT_ROWS_PER_BLOCK, n_rows);
  const float scale = (c < C) ? ldexpf(1.0f, scale_exponent(amax_bits[c])) : 0.0f;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
    uint32_t w = 0;
    if (c < C) {
      const float xs = src[(row0 + r) * ld + c] * scale;
      const __half hi = __float2half_rn(xs);
      const __half lo = __float2half_rn(xs - __half2float(hi));
      w = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
    }
    dst[r * ldp + c] = w;
  }
}

// inv[r] = 1 / scale of the channel behind tile-space row r (0 for padding rows).  Column j of the
// design belongs to stream 0 below col1 (channel j % C0), else to stream 1 (channel (j - col1) % C1).
__global__ void row_inv_scale_kernel(const int* __restrict__ out_idx, int64_t n, int64_t col1, int C0, int C1,
                                     const unsigned* __restrict__ amax0, const unsigned* __restrict__ amax1,
                                     float* __restrict__ inv) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int j = out_idx[r];
  float v = 0.0f;
  if (j >= 0) {
    const bool second = C1 > 0 && j >= col1;
    const unsigned bits = second ? amax1[(j - col1) % C1] : amax0[j % C0];
    v = ldexpf(1.0f, -scale_exponent(bits));
  }
  inv[r] = v;
}

// ---------------------------------------------------------------------------------------------
// Host-side plan.
// ---------------------------------------------------------------------------------------------
struct GramPlan {
  int n_streams = 0, CB = 0, cg = 1, n_seg = 0, n_shift = 4;
  int L[2] = {0, 0}, C[2] = {0, 0}, Cpad[2] = {0, 0}, pre[2] = {0, 0};
  int64_t Tp = 0, Dz = 0, Dpad = 0, NK = 0;
  int nt = 0, ntri = 0, n_groups = 0, tile = 0;
  std::vector<int64_t> seg_off;
  size_t off_stage[2] = {0, 0}, off_groups = 0, off_ktiles = 0, off_outidx = 0, off_segoff = 0,
         off_segtab = 0, total = 0;
};

int pick_cb(int channels_min_max) {
  // Largest TMA box height whose channel padding wastes <= 10 %.
  const int cands[4] = {64, 32, 16, 8};
  for (int cb : cands) {
    const int64_t padded = round_up(channels_min_max, cb);
    if ((double)padded <= 1.10 * channels_min_max) return cb;
  }
  return 8;
}

inline int64_t first_of(const int64_t* seg_first, int s) { return seg_first ? seg_first[s] : 0; }

int make_plan(const td_stream* streams, int n_streams, const int64_t* seg_first,
              const int64_t* seg_keep, int n_seg, const td_gram_options* opt, GramPlan* pl) {
  TD_REQUIRE(n_streams == 1 || n_streams == 2, "td_lagged_gram: n_streams must be 1 or 2, not %d",
             n_streams);
  TD_REQUIRE(n_seg > 0, "td_lagged_gram: no segments");
  for (int q = 0; q < n_streams; ++q) {
    int rc = check_stream(&streams[q], seg_keep, n_seg, "td_lagged_gram");
    if (rc) return rc;
  }
  pl->n_streams = n_streams;
  pl->n_seg = n_seg;
  pl->n_shift = 4;
  int cg = opt ? opt->cta_group : 0;
  int cb = 64;
  int halo = 0;
  pl->Dz = 0;
  for (int q = 0; q < n_streams; ++q) {
    pl->L[q] = lags_of(streams[q]);
    pl->C[q] = streams[q].channels;
    pl->pre[q] = streams[q].pre;
    cb = std::min(cb, pick_cb(streams[q].channels));
    halo = std::max(halo, streams[q].pre + streams[q].post);
    pl->Dz += width_of(streams[q]);
  }
  pl->CB = cb;
  int64_t n_groups = 0;
  for (int q = 0; q < n_streams; ++q) {
    pl->Cpad[q] = (int)round_up(pl->C[q], cb);
    n_groups += (int64_t)pl->L[q] * (pl->Cpad[q] / cb);
  }
  // CTA pairs (256x256 tiles, A operand in tensor memory) from two pair tiles per side on: measured
  // at D = 512 (64 channels x 8 lags, 20 000 frames) 63 vs 44 TFLOP/s for single-CTA 128x128 tiles.
  if (cg == 0) cg = (n_groups * cb >= 512) ? 2 : 1;
  TD_REQUIRE(cg == 1 || cg == 2, "td_lagged_gram: cta_group must be 0, 1 or 2");
  pl->cg = cg;
  pl->tile = TILE_ROWS * cg;
  pl->Dpad = round_up(n_groups * cb, pl->tile);
  pl->nt = (int)(pl->Dpad / pl->tile);
  pl->n_groups = (int)(pl->Dpad / cb);
  pl->ntri = pl->nt * (pl->nt + 1) / 2;
  TD_REQUIRE(pl->Dpad < (1LL << 30), "td_lagged_gram: design too wide");

  // Segment placement in staged time: zero gap >= halo between segments, 32-frame aligned starts.
  pl->seg_off.resize(n_seg);
  int64_t cur = round_up(halo, 32);
  int64_t nk = 0;
  for (int s = 0; s < n_seg; ++s) {
    pl->seg_off[s] = cur;
    int64_t len = 0;
    for (int q = 0; q < n_streams; ++q) len = std::max(len, streams[q].seg_len[s]);
    cur = round_up(cur + len + halo, 32);
    const int64_t f = first_of(seg_first, s);
    TD_REQUIRE(f >= 0, "td_lagged_gram: negative first row at segment %d", s);
    if (seg_keep[s] > f) nk += ceil_div(seg_keep[s], BK) - f / BK;
  }
  pl->Tp = round_up(cur + BK + halo, 32);
  pl->NK = nk;
  TD_REQUIRE(pl->Tp < (1LL << 31) - 4096, "td_lagged_gram: more than 2^31 staged frames");

  size_t off = 0;
  for (int q = 0; q < n_streams; ++q) {
    pl->off_stage[q] = off;
    off += (size_t)pl->n_shift * pl->Cpad[q] * pl->Tp * sizeof(float);
    off = (size_t)round_up((int64_t)off, 1024);
  }
  pl->off_groups = off;  off += (size_t)round_up((int64_t)pl->n_groups * sizeof(int4), 1024);
  pl->off_ktiles = off;  off += (size_t)round_up(std::max<int64_t>(nk, 1) * sizeof(int2), 1024);
  pl->off_outidx = off;  off += (size_t)round_up(pl->Dpad * sizeof(int), 1024);
  pl->off_segoff = off;  off += (size_t)round_up((int64_t)n_seg * sizeof(int64_t), 1024);
  pl->off_segtab = off;  off += (size_t)round_up((int64_t)n_seg * 2 * n_streams * sizeof(int64_t), 1024);
  pl->total = off;
  return TD_OK;
}

// TD_DEBUG=1: synchronise after every phase and report on stderr (development aid).
bool debug_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TD_DEBUG"); v = (e && e[0] && e[0] != '0') ? 1 : 0; }
  return v == 1;
}
int debug_sync(cudaStream_t st, const char* what) {
  if (!debug_enabled()) return TD_OK;
  cudaError_t e = cudaStreamSynchronize(st);
  fprintf(stderr, "[tdgram] %s: %s\n", what, cudaGetErrorString(e));
  fflush(stderr);
  if (e != cudaSuccess) { set_error("%s failed: %s", what, cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
thread_local bool g_ev_valid = false;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TD_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return TD_ERR_CUDA;
    }
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return TD_OK;
}

int make_tmap(CUtensorMap* m, float* planes, int64_t Tp, int Cpad, int CB, int n_shift) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  cuuint64_t dims[3] = {(cuuint64_t)Tp, (cuuint64_t)Cpad, (cuuint64_t)n_shift};
  cuuint64_t strides[2] = {(cuuint64_t)Tp * 4, (cuuint64_t)Tp * 4 * (cuuint64_t)Cpad};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)CB, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, planes, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (Tp=%lld Cpad=%d CB=%d)", (int)r,
              (long long)Tp, Cpad, CB);
    return TD_ERR_CUDA;
  }
  return TD_OK;
}


// ---------------------------------------------------------------------------------------------
// Halo kernel (method 4): plan, tables, launch.
// ---------------------------------------------------------------------------------------------
struct HaloPlan {
  int n_streams = 0, n_seg = 0;
  int L[2] = {0, 0}, C[2] = {0, 0}, pre[2] = {0, 0}, post[2] = {0, 0}, ncg[2] = {0, 0}, nlg[2] = {0, 0};
  int64_t Dz = 0, Dpad = 0, NK = 0;
  int nt = 0, ntri = 0, n_rg = 0;
  bool staged[2] = {false, false};
  bool f16 = false;
  int64_t ld_eff[2] = {0, 0}, row_lo[2] = {0, 0}, n_rows[2] = {0, 0};
  size_t off_stage[2] = {0, 0}, off_rgs = 0, off_ktiles = 0, off_outidx = 0, off_amax[2] = {0, 0}, off_inv = 0,
         total = 0;
};

int make_halo_plan(const td_stream* streams, int n_streams, const int64_t* seg_first,
                   const int64_t* seg_keep, int n_seg, HaloPlan* pl, bool f16 = false) {
  pl->f16 = f16;
  TD_REQUIRE(n_streams == 1 || n_streams == 2, "td_lagged_gram: n_streams must be 1 or 2, not %d", n_streams);
  TD_REQUIRE(n_seg > 0, "td_lagged_gram: no segments");
  pl->n_streams = n_streams;
  pl->n_seg = n_seg;
  int64_t n_rg = 0;
  for (int q = 0; q < n_streams; ++q) {
    int rc = check_stream(&streams[q], seg_keep, n_seg, "td_lagged_gram");
    if (rc) return rc;
    pl->L[q] = lags_of(streams[q]);
    pl->C[q] = streams[q].channels;
    pl->pre[q] = streams[q].pre;
    pl->post[q] = streams[q].post;
    pl->ncg[q] = (int)ceil_div(pl->C[q], hk::BOX_CH);
    pl->nlg[q] = (int)ceil_div(pl->L[q], hk::RG_LAGS);
    n_rg += (int64_t)pl->ncg[q] * pl->nlg[q];
    pl->Dz += width_of(streams[q]);
  }
  pl->n_rg = (int)n_rg;
  pl->Dpad = round_up(n_rg * hk::RG_ROWS, 2 * TILE_ROWS);
  pl->nt = (int)(pl->Dpad / (2 * TILE_ROWS));
  pl->ntri = pl->nt * (pl->nt + 1) / 2;
  TD_REQUIRE(pl->Dpad < (1LL << 30), "td_lagged_gram: design too wide");
  int64_t nk = 0;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t f = first_of(seg_first, s);
    TD_REQUIRE(f >= 0, "td_lagged_gram: negative first row at segment %d", s);
    if (seg_keep[s] > f) nk += ceil_div(seg_keep[s], BK) - f / BK;
  }
  pl->NK = nk;
  size_t off = 0;
  for (int q = 0; q < n_streams; ++q) {
    // Raw rows this call can touch: design rows [first, keep) of every used segment +- the lag halo.
    int64_t lo = -1, hi = 0, extent = 0;
    for (int s = 0; s < n_seg; ++s) {
      extent = std::max(extent, streams[q].seg_start[s] + streams[q].seg_len[s]);
      const int64_t f = first_of(seg_first, s);
      if (seg_keep[s] <= f) continue;
      const int64_t a = streams[q].seg_start[s] + std::max<int64_t>(f - pl->pre[q], 0);
      const int64_t b = streams[q].seg_start[s] + std::min<int64_t>(seg_keep[s] + pl->post[q], streams[q].seg_len[s]);
      lo = (lo < 0) ? a : std::min(lo, a);
      hi = std::max(hi, b);
    }
    if (lo < 0) lo = hi = 0;
    // TMA needs a 16-byte aligned base and row stride; otherwise the touched rows are copied once
    // into a padded buffer (1x the touched part of the recording, not the 4 shifted planes of the
    // earlier kernels).
    const bool aligned = (streams[q].ld % 4 == 0) && (reinterpret_cast<uintptr_t>(streams[q].data_dev) % 16 == 0);
    pl->staged[q] = !aligned || f16;                 // fp16 x 3 always reads its pre-split copy
    if (!pl->staged[q]) {
      pl->ld_eff[q] = streams[q].ld;
      pl->row_lo[q] = 0;
      pl->n_rows[q] = extent;
    } else {
      pl->ld_eff[q] = round_up(pl->C[q], 4);
      pl->row_lo[q] = lo;
      pl->n_rows[q] = hi - lo;
      pl->off_stage[q] = off;
      off += (size_t)round_up(std::max<int64_t>(pl->n_rows[q], 1) * pl->ld_eff[q] * (int64_t)sizeof(float), 1024);
    }
    TD_REQUIRE(pl->n_rows[q] < (1LL << 31) - 4096, "td_lagged_gram: more than 2^31 raw rows");
  }
  pl->off_rgs = off;     off += (size_t)round_up((pl->Dpad / hk::RG_ROWS) * (int64_t)sizeof(int4), 1024);
  pl->off_ktiles = off;  off += (size_t)round_up(std::max<int64_t>(nk, 1) * 2 * (int64_t)sizeof(int4), 1024);
  pl->off_outidx = off;  off += (size_t)round_up(pl->Dpad * (int64_t)sizeof(int), 1024);
  if (f16) {
    for (int q = 0; q < n_streams; ++q) {
      pl->off_amax[q] = off;
      off += (size_t)round_up(pl->C[q] * (int64_t)sizeof(unsigned), 1024);
    }
    pl->off_inv = off;   off += (size_t)round_up(pl->Dpad * (int64_t)sizeof(float), 1024);
  }
  pl->total = off;
  return TD_OK;
}

// The staged copy of stream q: a padded fp32 copy (rows TMA cannot address), or, for the fp16 x 3
// flavour, the pre-split copy (absmax pass + split pass; amax stays in the workspace for the epilogue's
// inverse scales).
int stage_stream(const HaloPlan& pl, const td_stream* streams, int q, uint8_t* ws, cudaStream_t st) {
  float* dst = reinterpret_cast<float*>(ws + pl.off_stage[q]);
  if (pl.n_rows[q] <= 0) {
    if (pl.f16) TD_CHECK_CUDA(cudaMemsetAsync(ws + pl.off_amax[q], 0, pl.C[q] * sizeof(unsigned), st));
    return TD_OK;
  }
  if (!pl.f16) {
    const int64_t total = pl.n_rows[q] * pl.ld_eff[q];
    pad_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(streams[q].data_dev, streams[q].ld, pl.C[q],
                                                                  pl.row_lo[q], pl.n_rows[q], dst, (int)pl.ld_eff[q]);
    TD_CHECK_CUDA(cudaGetLastError());
    return TD_OK;
  }
  unsigned* amax = reinterpret_cast<unsigned*>(ws + pl.off_amax[q]);
  TD_CHECK_CUDA(cudaMemsetAsync(amax, 0, pl.C[q] * sizeof(unsigned), st));
  const dim3 block(32, 8);
  const int n_cblocks = (int)ceil_div(pl.ld_eff[q], 32);
  const int64_t n_blocks = ceil_div(pl.n_rows[q], SPLIT_ROWS_PER_BLOCK) * n_cblocks;
  TD_REQUIRE(n_blocks < (1LL << 31), "td_lagged_gram: recording too large for the split pass");
  absmax_kernel<<<(unsigned)n_blocks, block, 0, st>>>(streams[q].data_dev, streams[q].ld, pl.C[q], pl.row_lo[q], pl.n_rows[q],
                                                      amax, n_cblocks);
  TD_CHECK_CUDA(cudaGetLastError());
  split_rows_kernel<<<(unsigned)n_blocks, block, 0, st>>>(streams[q].data_dev, streams[q].ld, pl.C[q], pl.row_lo[q], pl.n_rows[q],
                                                          amax, reinterpret_cast<uint32_t*>(dst), (int)pl.ld_eff[q], n_cblocks);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}

int make_tmap_raw(CUtensorMap* m, const float* base, int64_t n_rows, int C, int64_t ld,
                  int box_ch = hk::BOX_CH, int box_frames = hk::BOX_FRAMES) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)std::max<int64_t>(n_rows, 1)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_ch, (cuuint32_t)box_frames};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (raw stream) failed with CUresult %d (rows=%lld C=%d ld=%lld)", (int)r,
              (long long)n_rows, C, (long long)ld);
    return TD_ERR_CUDA;
  }
  return TD_OK;
}

// Which flavour of the halo kernel a call gets: fp16 x 3 (method 6) unless 3xTF32 (method 4) is asked
// for by number.  Same split arithmetic, 1.4 - 1.5x the frames per joule on a power-capped B200.
inline bool halo_is_f16(const td_gram_options* opt) { return !opt || opt->method != 4; }

// batched: one Gram per segment, G [n_seg][Dz][ldg]; every segment must span the same number of
// k-tiles (then a superchunk is exactly one segment and the epilogue picks the output by superchunk).
int run_halo(const td_stream* streams, int n_streams, const int64_t* seg_first, const int64_t* seg_keep,
             int n_seg, float* G, int64_t ldg, const td_gram_options* opt, void* workspace,
             size_t workspace_bytes, cudaStream_t st, bool batched = false) {
  HaloPlan pl;
  const bool f16 = halo_is_f16(opt);
  int rc = make_halo_plan(streams, n_streams, seg_first, seg_keep, n_seg, &pl, f16);
  if (rc) return rc;
  int64_t kt_per_seg = 0;
  if (batched) {
    for (int s = 0; s < n_seg; ++s) {
      const int64_t f = first_of(seg_first, s);
      const int64_t k = seg_keep[s] > f ? ceil_div(seg_keep[s], BK) - f / BK : 0;
      if (s == 0) kt_per_seg = k;
      if (k != kt_per_seg || k == 0) {
        set_error("td_lagged_gram_batched_f32: segment %d spans %lld k-tiles, segment 0 spans %lld (one launch needs equal, non-empty segments)",
                  s, (long long)k, (long long)kt_per_seg);
        return TD_ERR_UNSUPPORTED;
      }
    }
  }
  TD_REQUIRE(ldg >= pl.Dz, "td_lagged_gram_f32: ldg (%lld) < Dz (%lld)", (long long)ldg, (long long)pl.Dz);
  if (workspace_bytes < pl.total || !workspace) {
    set_error("td_lagged_gram_f32: workspace has %zu bytes, %zu needed", workspace_bytes, pl.total);
    return TD_ERR_WORKSPACE;
  }
  int dev = 0, sms = 0, cc_major = 0;
  TD_CHECK_CUDA(cudaGetDevice(&dev));
  TD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TD_CHECK_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) {
    set_error("td_lagged_gram_f32: the tcgen05 kernel needs an sm_100 device (found sm_%d)", cc_major * 10);
    return TD_ERR_UNSUPPORTED;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  // ---- row groups: (stream, channel group, lag group), lag group fastest, so that the row groups of
  // a 128-row operand tile are mostly consecutive lags of the same 8 channels and share one raw box.
  const int n_rg_pad = (int)(pl.Dpad / hk::RG_ROWS);
  std::vector<int4> rgs(n_rg_pad, make_int4(0, 0, 0, 0));
  std::vector<int> out_idx(pl.Dpad, -1);
  {
    struct Key { int q, cg, lg; };
    std::vector<Key> keys;
    keys.reserve(pl.n_rg);
    for (int q = 0; q < n_streams; ++q)
      for (int cg = 0; cg < pl.ncg[q]; ++cg)
        for (int lg = 0; lg < pl.nlg[q]; ++lg) keys.push_back({q, cg, lg});
    int64_t col0[2] = {0, n_streams == 2 ? width_of(streams[0]) : 0};
    int head = -1;                                        // index of the run head inside the CTA tile
    for (int i = 0; i < pl.n_rg; ++i) {
      const Key& k = keys[i];
      const int r = i % hk::RG_PER_CTA;
      const bool follows = r > 0 && head >= 0 && keys[i - 1].q == k.q && keys[i - 1].cg == k.cg &&
                           keys[i - 1].lg + 1 == k.lg;
      if (!follows) head = i;
      const int head_r = head % hk::RG_PER_CTA;
      const int off = head_r * hk::BOX_BYTES + (k.lg - keys[head].lg) * hk::RG_LAGS * hk::BOX_CH * 4;
      rgs[i] = make_int4(k.q | ((follows ? 0 : 1) << 1), k.cg * hk::BOX_CH,
                         k.lg * hk::RG_LAGS - pl.pre[k.q], off);
      for (int kk = 0; kk < hk::RG_LAGS; ++kk)
        for (int cc = 0; cc < hk::BOX_CH; ++cc) {
          const int lag = k.lg * hk::RG_LAGS + kk, c = k.cg * hk::BOX_CH + cc;
          if (lag < pl.L[k.q] && c < pl.C[k.q])
            out_idx[(int64_t)i * hk::RG_ROWS + kk * hk::BOX_CH + cc] = (int)(col0[k.q] + (int64_t)lag * pl.C[k.q] + c);
        }
      if (r == hk::RG_PER_CTA - 1) head = -1;
    }
  }
  std::vector<int4> ktiles;
  ktiles.reserve((size_t)pl.NK * 2);
  for (int s = 0; s < n_seg; ++s) {
    const int64_t f = first_of(seg_first, s);
    for (int64_t j = f / BK; j * BK < seg_keep[s] && seg_keep[s] > f; ++j) {
      const int64_t t0 = j * BK;
      const int begin = (int)std::max<int64_t>(0, f - t0);
      const int end = (int)std::min<int64_t>(BK, seg_keep[s] - t0);
      int row0[2] = {0, 0}, hi[2] = {0, 0};
      int edge = 0;
      for (int q = 0; q < n_streams; ++q) {
        row0[q] = (int)(streams[q].seg_start[s] + t0 - pl.row_lo[q]);
        const int64_t len = streams[q].seg_len[s];
        hi[q] = (int)std::max<int64_t>(std::min<int64_t>(len - t0, 1 << 24), -(1 << 24));
        if (t0 - pl.pre[q] < 0 || t0 + BK - 1 + pl.post[q] >= len) edge = 1;
      }
      const int neg_t0 = (int)std::max<int64_t>(-t0, -(1 << 24));
      ktiles.push_back(make_int4(row0[0], row0[1], begin | (end << 8) | (edge << 16), neg_t0));
      ktiles.push_back(make_int4(hi[0], hi[1], 0, 0));
    }
  }
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_rgs, rgs.data(), rgs.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
  if (!ktiles.empty())
    TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_ktiles, ktiles.data(), ktiles.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_outidx, out_idx.data(), out_idx.size() * sizeof(int), cudaMemcpyHostToDevice, st));

  // ---- tensor maps over the raw streams (a padded copy only where TMA cannot address the rows) ----
  CUtensorMap tmaps[2];
  for (int q = 0; q < n_streams; ++q) {
    const float* base = streams[q].data_dev;
    if (pl.staged[q]) {
      if ((rc = stage_stream(pl, streams, q, ws, st))) return rc;
      base = reinterpret_cast<const float*>(ws + pl.off_stage[q]);
    }
    // (the pre-split copy has ld_eff words per row, channels up to ld_eff are real zeros)
    rc = make_tmap_raw(&tmaps[q], base, pl.n_rows[q], f16 ? (int)pl.ld_eff[q] : pl.C[q], pl.ld_eff[q]);
    if (rc) return rc;
  }
  if (n_streams == 1) tmaps[1] = tmaps[0];
  if (f16) {
    row_inv_scale_kernel<<<(unsigned)ceil_div(pl.Dpad, 256), 256, 0, st>>>(
        reinterpret_cast<const int*>(ws + pl.off_outidx), pl.Dpad, n_streams == 2 ? width_of(streams[0]) : 0, pl.C[0],
        n_streams == 2 ? pl.C[1] : 0, reinterpret_cast<const unsigned*>(ws + pl.off_amax[0]),
        reinterpret_cast<const unsigned*>(ws + pl.off_amax[n_streams == 2 ? 1 : 0]), reinterpret_cast<float*>(ws + pl.off_inv));
    TD_CHECK_CUDA(cudaGetLastError());
  }
  if ((rc = debug_sync(st, "halo stage"))) return rc;

  const int64_t n_out = batched ? n_seg : 1;
  TD_CHECK_CUDA(cudaMemset2DAsync(G, (size_t)ldg * sizeof(float), 0, (size_t)pl.Dz * sizeof(float), (size_t)(pl.Dz * n_out), st));
  if (pl.NK > 0) {
    hk::HaloParams P;
    P.batch_stride = batched ? pl.Dz * ldg : 0;
    P.rect_ta = 0;
    P.rgs_a = nullptr;
    P.out_idx_a = nullptr;
    P.rgs = reinterpret_cast<const int4*>(ws + pl.off_rgs);
    P.ktiles = reinterpret_cast<const int4*>(ws + pl.off_ktiles);
    P.out_idx = reinterpret_cast<const int*>(ws + pl.off_outidx);
    P.inv_scale = f16 ? reinterpret_cast<const float*>(ws + pl.off_inv) : nullptr;
    P.inv_scale_a = nullptr;
    P.G = G;
    P.ldg = ldg;
    P.ntri = pl.ntri;
    P.NK = pl.NK;
    // TMEM accumulates chunk_rows frames in fp32 before the epilogue folds them into its registers: the
    // rounding error grows with the chunk (fp16 x 3: 7e-7 / 1.4e-6 / 2.9e-6 relative Frobenius error at
    // 128 / 256 / 512 rows, 3xTF32: 2.3e-6 at 128), the drain cost shrinks with it.
    const int chunk_rows = (opt && opt->chunk_rows > 0) ? opt->chunk_rows : (f16 ? 256 : 128);
    P.chunk_kt = std::max(1, chunk_rows / BK);
    // Superchunk: the raw frames all tiles sweep together; sized to stay L2 resident (~48 MB of the
    // 126 MB: measured on the cfg 4 width, 24 / 48 / 96 / 192 MB -> 1077 / 1072 / 1088 / 1103 ms).
    int64_t sc_rows = (opt && opt->superchunk_rows > 0) ? opt->superchunk_rows : 0;
    if (sc_rows == 0) {
      int64_t row_bytes = 0;
      for (int q = 0; q < n_streams; ++q) row_bytes += pl.ld_eff[q] * 4;
      sc_rows = std::max<int64_t>((48LL << 20) / row_bytes, 4096);
    }
    int64_t ks = std::max<int64_t>(P.chunk_kt, round_up(ceil_div(sc_rows, BK), P.chunk_kt));
    ks = std::min<int64_t>(ks, round_up(pl.NK, P.chunk_kt));
    // Equal superchunks: a short last one would leave part of the workers idle for its whole sweep.
    int64_t nsc = ceil_div(pl.NK, ks);
    ks = round_up(ceil_div(pl.NK, nsc), P.chunk_kt);
    nsc = ceil_div(pl.NK, ks);
    if (batched) { ks = kt_per_seg; nsc = n_seg; }
    P.KS = (int)ks;
    P.total_steps = nsc * pl.ntri * ks;
    int workers = sms / 2;
    if (opt && opt->max_ctas > 0) workers = std::max(1, std::min(workers, opt->max_ctas / 2));
    workers = (int)std::max<int64_t>(1, std::min<int64_t>(workers, ceil_div((int64_t)pl.ntri * ks, P.chunk_kt)));
    P.prof = nullptr;
    const bool timed = opt && opt->reserved[0] == 1;
    if (timed) {
      if (!g_ev_start) {
        TD_CHECK_CUDA(cudaEventCreate(&g_ev_start));
        TD_CHECK_CUDA(cudaEventCreate(&g_ev_stop));
      }
      TD_CHECK_CUDA(cudaEventRecord(g_ev_start, st));
    }
    auto kernel = f16 ? hk::lagged_gram_halo_kernel<false, true> : hk::lagged_gram_halo_kernel<false, false>;
    const size_t smem_bytes = f16 ? hk::F16_SMEM_BYTES : hk::SMEM_BYTES;
    TD_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * 2);
    cfg.blockDim = dim3(f16 ? hk::NUM_THREADS_F16 : hk::NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmaps[0], tmaps[1], P));
    TD_CHECK_CUDA(cudaGetLastError());
    if (timed) {
      TD_CHECK_CUDA(cudaEventRecord(g_ev_stop, st));
      g_ev_valid = true;
    }
    if (debug_enabled())
      fprintf(stderr, "[tdgram] halo launch: workers=%d Dz=%lld Dpad=%lld ntri=%d NK=%lld KS=%d chunk_kt=%d staged=%d/%d\n",
              workers, (long long)pl.Dz, (long long)pl.Dpad, pl.ntri, (long long)pl.NK, P.KS, P.chunk_kt,
              (int)pl.staged[0], (int)pl.staged[1]);
    if ((rc = debug_sync(st, "halo gram kernel"))) return rc;
  }
  const int nb = (int)ceil_div(pl.Dz, 32);
  for (int64_t o0 = 0; o0 < n_out; o0 += 32768) {
    const unsigned nz = (unsigned)std::min<int64_t>(32768, n_out - o0);
    fold_symmetric_kernel<<<dim3(nb, nb, nz), dim3(32, 8), 0, st>>>(G + o0 * pl.Dz * ldg, pl.Dz, ldg, pl.Dz * ldg);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  return TD_OK;
}

// ---------------------------------------------------------------------------------------------
// Method 5: the Gram of a lagged design is block-Toeplitz up to edge terms.
//
//   G[(k1,c1),(k2,c2)] = sum_{t in rows} x~[t-pre+k1, c1] x~[t-pre+k2, c2]      (x~ = x, zero outside the trial)
//                      = C(d)[c1,c2] + sum_seg ( T_seg(k1, d) - H_seg(k1, d) ),    d = k2 - k1 >= 0
//   C(d)[c1,c2] = sum_{t in rows} x~[t-pre, c1] x~[t-pre+d, c2]                  (the k1 = 0 block row)
//   H(k1, d) = sum_{u = first-pre}^{first-pre+k1-1} x~[u,c1] x~[u+d,c2]           (k1 frames at the head)
//   T(k1, d) = sum_{u = keep-pre}^{keep-pre+k1-1}  x~[u,c1] x~[u+d,c2]           (k1 frames past the tail)
//
// So only the first block row is a real contraction -- C x (C L) outputs instead of (C L)^2 / 2: the
// tcgen05 kernel runs in rectangular mode (A = the lag-0 rows, B = every row), L / 2 times fewer MMAs
// -- and toeplitz_expand_kernel fills G along the lag diagonals, carrying the edge sums (k1 more
// outer products per diagonal step and segment).  Same 3xTF32 arithmetic, same row semantics
// (seg_first / seg_keep, zero padding at trial edges); one stream only.
// ---------------------------------------------------------------------------------------------
struct SegEdge { long long start, len, first, keep; };

// grid: (n_batch * ceil(C / 32), C, L): thread = (c2, c1, d).  Ct[(d*C + c2)][c1] holds C(d)[c1, c2].
// Batched (one Gram per segment): batch b reads Ct + b*ct_stride, segment b only, writes G + b*g_stride.
__global__ void __launch_bounds__(32) toeplitz_expand_kernel(const float* __restrict__ Ct, long long ldc,
                                                             const float* __restrict__ x, long long ld, int C, int L,
                                                             int pre, const SegEdge* __restrict__ segs, int n_seg,
                                                             float* __restrict__ G, long long ldg, int cblocks,
                                                             long long ct_stride, long long g_stride) {
  const int batch = blockIdx.x / cblocks;
  const int c2 = (blockIdx.x - batch * cblocks) * 32 + threadIdx.x;
  const int c1 = blockIdx.y;
  const int d = blockIdx.z;
  if (g_stride) {                              // one segment per output
    Ct += (long long)batch * ct_stride;
    G += (long long)batch * g_stride;
    segs += batch;
    n_seg = 1;
  }
  // Upper triangle only (i <= j in the design's lag-major column order; td_symmetrize_f32 mirrors it):
  // on the block diagonal (d = 0) that is c1 <= c2, so every entry has exactly one writer.
  if (c2 >= C || (d == 0 && c2 < c1)) return;
  const double base = (double)Ct[((long long)d * C + c2) * ldc + c1];
  double run = 0.0;
  for (int k1 = 0; k1 + d < L; ++k1) {
    const float v = (float)(base + run);
    const long long i = (long long)k1 * C + c1, j = (long long)(k1 + d) * C + c2;
    G[i * ldg + j] = v;
    // edge terms of the next diagonal step: one more frame at the tail, one fewer at the head
    double delta = 0.0;
    for (int s = 0; s < n_seg; ++s) {
      const SegEdge e = segs[s];
      if (e.keep <= e.first) continue;
      const float* xs = x + e.start * ld;
      const long long ut = e.keep - pre + k1, uh = e.first - pre + k1;
      if (ut >= 0 && ut + d < e.len) delta += (double)__ldg(xs + ut * ld + c1) * (double)__ldg(xs + (ut + d) * ld + c2);
      if (uh >= 0 && uh + d < e.len) delta -= (double)__ldg(xs + uh * ld + c1) * (double)__ldg(xs + (uh + d) * ld + c2);
    }
    run += delta;
  }
}

int run_toeplitz(const td_stream* streams, int n_streams, const int64_t* seg_first, const int64_t* seg_keep,
                 int n_seg, float* G, int64_t ldg, const td_gram_options* opt, void* workspace,
                 size_t workspace_bytes, cudaStream_t st, size_t* bytes_only, bool batched = false);

// Does this call run the halo kernel (methods 4 / 6)?  Explicitly, or automatically (method 0: the
// fp16 x 3 flavour, see halo_is_f16) when the design is wide enough for CTA pairs (>= 512 tile rows).
constexpr int DEFAULT_PAIR_METHOD = 4;
int use_halo(const td_stream* streams, int n_streams, const td_gram_options* opt, bool* halo) {
  const int m = opt ? opt->method : 0;
  const int cg = opt ? opt->cta_group : 0;
  TD_REQUIRE(m >= 0 && m <= 6, "td_lagged_gram: unknown method %d", m);
  TD_REQUIRE(cg >= 0 && cg <= 2, "td_lagged_gram: cta_group must be 0, 1 or 2");
  TD_REQUIRE(!((m == 4 || m == 6) && cg == 1), "td_lagged_gram: methods 4 and 6 (operand built on chip) need CTA pairs");
  *halo = (m == 4 || m == 6);
  if (m == 0 && DEFAULT_PAIR_METHOD == 4 && cg != 1 && streams) {
    // Row groups are 4 lags x 8 channels: a design with very few lags (or channels) would mostly
    // multiply padding (one lag fills a quarter of its row groups); those go to the staged kernels.
    int64_t rows = 0, cols = 0;
    for (int q = 0; q < n_streams && q < 2; ++q) {
      rows += ceil_div(streams[q].channels, hk::BOX_CH) * hk::BOX_CH * ceil_div(lags_of(streams[q]), hk::RG_LAGS) * hk::RG_LAGS;
      cols += width_of(streams[q]);
    }
    *halo = (cg == 2 || rows >= 512) && cols * 5 >= rows * 4;
  }
  return TD_OK;
}


// Workspace layout of method 5: [halo plan (tables, optional padded copy)] [A row groups] [A out_idx]
// [segment edges] [Ct: Dz x ldc floats].  bytes_only != nullptr: just report the size.
int run_toeplitz(const td_stream* streams, int n_streams, const int64_t* seg_first, const int64_t* seg_keep,
                 int n_seg, float* G, int64_t ldg, const td_gram_options* opt, void* workspace,
                 size_t workspace_bytes, cudaStream_t st, size_t* bytes_only, bool batched) {
  if (n_streams != 1) {
    set_error("td_lagged_gram: method 5 (block-Toeplitz structure) handles one stream");
    return TD_ERR_UNSUPPORTED;
  }
  HaloPlan pl;
  // The block row runs on the default dense flavour (fp16 x 3 on the pre-split copy) unless
  // reserved[1] = 1 asks for 3xTF32 on the raw array.
  const bool f16 = !(opt && opt->reserved[1] == 1);
  int rc = make_halo_plan(streams, n_streams, seg_first, seg_keep, n_seg, &pl, f16);
  if (rc) return rc;
  const int C = pl.C[0], L = pl.L[0], pre = pl.pre[0];
  const int n_rga = (int)ceil_div(C, hk::ABOX_CH);                    // A row groups: 32 channels at lag 0
  const int n_rga_pad = (int)round_up(n_rga, 2 * hk::RG_PER_CTA);     // whole pair tiles of A rows
  const int ta = n_rga_pad / (2 * hk::RG_PER_CTA);
  const int64_t ldc = round_up(C, 4);
  size_t off = pl.total;
  const size_t off_rga = off;   off += (size_t)round_up((int64_t)n_rga_pad * (int64_t)sizeof(int4), 1024);
  const size_t off_oia = off;   off += (size_t)round_up((int64_t)n_rga_pad * hk::RG_ROWS * (int64_t)sizeof(int), 1024);
  const size_t off_seg = off;   off += (size_t)round_up((int64_t)n_seg * (int64_t)sizeof(SegEdge), 1024);
  const size_t off_inva = off;  off += (size_t)round_up((int64_t)n_rga_pad * hk::RG_ROWS * (int64_t)sizeof(float), 1024);
  const int64_t n_out = batched ? n_seg : 1;
  const size_t off_ct = off;    off += (size_t)round_up(n_out * pl.Dz * ldc * (int64_t)sizeof(float), 1024);
  if (bytes_only) { *bytes_only = off; return TD_OK; }
  int64_t kt_per_seg = 0;
  if (batched) {
    for (int s = 0; s < n_seg; ++s) {
      const int64_t f = first_of(seg_first, s);
      const int64_t k = seg_keep[s] > f ? ceil_div(seg_keep[s], BK) - f / BK : 0;
      if (s == 0) kt_per_seg = k;
      if (k != kt_per_seg || k == 0) {
        set_error("td_lagged_gram_batched_f32: segment %d spans %lld k-tiles, segment 0 spans %lld (one launch needs equal, non-empty segments)",
                  s, (long long)k, (long long)kt_per_seg);
        return TD_ERR_UNSUPPORTED;
      }
    }
  }
  TD_REQUIRE(ldg >= pl.Dz, "td_lagged_gram_f32: ldg (%lld) < Dz (%lld)", (long long)ldg, (long long)pl.Dz);
  if (workspace_bytes < off || !workspace) {
    set_error("td_lagged_gram_f32: workspace has %zu bytes, %zu needed", workspace_bytes, off);
    return TD_ERR_WORKSPACE;
  }
  int dev = 0, sms = 0, cc_major = 0;
  TD_CHECK_CUDA(cudaGetDevice(&dev));
  TD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TD_CHECK_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) {
    set_error("td_lagged_gram_f32: the tcgen05 kernel needs an sm_100 device (found sm_%d)", cc_major * 10);
    return TD_ERR_UNSUPPORTED;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  // ---- B side: every row group, as in run_halo ------------------------------------------------
  const int n_rg_pad = (int)(pl.Dpad / hk::RG_ROWS);
  std::vector<int4> rgs(n_rg_pad, make_int4(0, 0, 0, 0));
  std::vector<int> out_idx(pl.Dpad, -1);
  {
    int head = -1, head_lg = 0;
    int i = 0;
    for (int cg = 0; cg < pl.ncg[0]; ++cg)
      for (int lg = 0; lg < pl.nlg[0]; ++lg, ++i) {
        const int r = i % hk::RG_PER_CTA;
        const bool follows = r > 0 && head >= 0 && lg > 0 && head_lg + (i - head) == lg;
        if (!follows) { head = i; head_lg = lg; }
        const int head_r = head % hk::RG_PER_CTA;
        rgs[i] = make_int4(0 | ((follows ? 0 : 1) << 1), cg * hk::BOX_CH, lg * hk::RG_LAGS - pre,
                           head_r * hk::BOX_BYTES + (lg - head_lg) * hk::RG_LAGS * hk::BOX_CH * 4);
        for (int kk = 0; kk < hk::RG_LAGS; ++kk)
          for (int cc = 0; cc < hk::BOX_CH; ++cc) {
            const int lag = lg * hk::RG_LAGS + kk, c = cg * hk::BOX_CH + cc;
            if (lag < L && c < C) out_idx[(int64_t)i * hk::RG_ROWS + kk * hk::BOX_CH + cc] = lag * C + c;
          }
        if (r == hk::RG_PER_CTA - 1) head = -1;
      }
  }
  // ---- A side: the lag-0 row of every channel, 32 channels per row group (its own 32 x 32 raw box) ----
  std::vector<int4> rga(n_rga_pad, make_int4(0, 0, 0, 0));
  std::vector<int> oia((size_t)n_rga_pad * hk::RG_ROWS, -1);
  for (int cg = 0; cg < n_rga; ++cg) {
    rga[cg] = make_int4(0 | (1 << 1), cg * hk::ABOX_CH, -pre, (cg % hk::RG_PER_CTA) * hk::ABOX_BYTES);
    for (int cc = 0; cc < hk::ABOX_CH; ++cc)
      if (cg * hk::ABOX_CH + cc < C) oia[(size_t)cg * hk::RG_ROWS + cc] = cg * hk::ABOX_CH + cc;
  }
  std::vector<int4> ktiles;
  ktiles.reserve((size_t)pl.NK * 2);
  std::vector<SegEdge> edges(n_seg);
  for (int s = 0; s < n_seg; ++s) {
    const int64_t f = first_of(seg_first, s);
    edges[s] = SegEdge{(long long)streams[0].seg_start[s], (long long)streams[0].seg_len[s], (long long)f,
                       (long long)seg_keep[s]};
    for (int64_t j = f / BK; j * BK < seg_keep[s] && seg_keep[s] > f; ++j) {
      const int64_t t0 = j * BK;
      const int begin = (int)std::max<int64_t>(0, f - t0);
      const int end = (int)std::min<int64_t>(BK, seg_keep[s] - t0);
      const int64_t len = streams[0].seg_len[s];
      const int row0 = (int)(streams[0].seg_start[s] + t0 - pl.row_lo[0]);
      const int hi = (int)std::max<int64_t>(std::min<int64_t>(len - t0, 1 << 24), -(1 << 24));
      const int edge = (t0 - pre < 0 || t0 + BK - 1 + pl.post[0] >= len) ? 1 : 0;
      ktiles.push_back(make_int4(row0, row0, begin | (end << 8) | (edge << 16), (int)std::max<int64_t>(-t0, -(1 << 24))));
      ktiles.push_back(make_int4(hi, hi, 0, 0));
    }
  }
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_rgs, rgs.data(), rgs.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
  if (!ktiles.empty())
    TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_ktiles, ktiles.data(), ktiles.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_outidx, out_idx.data(), out_idx.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + off_rga, rga.data(), rga.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + off_oia, oia.data(), oia.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + off_seg, edges.data(), edges.size() * sizeof(SegEdge), cudaMemcpyHostToDevice, st));

  const float* base = streams[0].data_dev;
  if (pl.staged[0]) {
    if ((rc = stage_stream(pl, streams, 0, ws, st))) return rc;
    base = reinterpret_cast<const float*>(ws + pl.off_stage[0]);
  }
  CUtensorMap tmap, tmap_a;
  const int c_map = f16 ? (int)pl.ld_eff[0] : C;      // (the pre-split copy has real zeros up to ld_eff)
  rc = make_tmap_raw(&tmap, base, pl.n_rows[0], c_map, pl.ld_eff[0]);
  if (rc) return rc;
  rc = make_tmap_raw(&tmap_a, base, pl.n_rows[0], c_map, pl.ld_eff[0], hk::ABOX_CH, hk::ABOX_FRAMES);
  if (rc) return rc;
  if (f16) {
    const unsigned* amax = reinterpret_cast<const unsigned*>(ws + pl.off_amax[0]);
    row_inv_scale_kernel<<<(unsigned)ceil_div(pl.Dpad, 256), 256, 0, st>>>(
        reinterpret_cast<const int*>(ws + pl.off_outidx), pl.Dpad, 0, C, 0, amax, amax, reinterpret_cast<float*>(ws + pl.off_inv));
    const int64_t n_a = (int64_t)n_rga_pad * hk::RG_ROWS;
    row_inv_scale_kernel<<<(unsigned)ceil_div(n_a, 256), 256, 0, st>>>(
        reinterpret_cast<const int*>(ws + off_oia), n_a, 0, C, 0, amax, amax, reinterpret_cast<float*>(ws + off_inva));
    TD_CHECK_CUDA(cudaGetLastError());
  }
  float* Ct = reinterpret_cast<float*>(ws + off_ct);
  TD_CHECK_CUDA(cudaMemsetAsync(Ct, 0, (size_t)n_out * pl.Dz * ldc * sizeof(float), st));
  if (pl.NK > 0) {
    hk::HaloParams P;
    P.rgs = reinterpret_cast<const int4*>(ws + pl.off_rgs);
    P.ktiles = reinterpret_cast<const int4*>(ws + pl.off_ktiles);
    P.out_idx = reinterpret_cast<const int*>(ws + pl.off_outidx);
    P.rect_ta = ta;
    P.inv_scale = f16 ? reinterpret_cast<const float*>(ws + pl.off_inv) : nullptr;
    P.inv_scale_a = f16 ? reinterpret_cast<const float*>(ws + off_inva) : nullptr;
    P.rgs_a = reinterpret_cast<const int4*>(ws + off_rga);
    P.out_idx_a = reinterpret_cast<const int*>(ws + off_oia);
    P.G = Ct;
    P.ldg = ldc;
    P.batch_stride = batched ? pl.Dz * ldc : 0;
    P.ntri = ta * pl.nt;
    P.NK = pl.NK;
    const int chunk_rows = (opt && opt->chunk_rows > 0) ? opt->chunk_rows : (f16 ? 256 : 128);
    P.chunk_kt = std::max(1, chunk_rows / BK);
    int64_t sc_rows = (opt && opt->superchunk_rows > 0) ? opt->superchunk_rows : 0;
    if (sc_rows == 0) sc_rows = std::max<int64_t>((48LL << 20) / (pl.ld_eff[0] * 4), 4096);
    int64_t ks = std::max<int64_t>(P.chunk_kt, round_up(ceil_div(sc_rows, BK), P.chunk_kt));
    ks = std::min<int64_t>(ks, round_up(pl.NK, P.chunk_kt));
    int64_t nsc = ceil_div(pl.NK, ks);
    ks = round_up(ceil_div(pl.NK, nsc), P.chunk_kt);
    nsc = ceil_div(pl.NK, ks);
    if (batched) { ks = kt_per_seg; nsc = n_seg; }
    P.KS = (int)ks;
    P.total_steps = nsc * P.ntri * ks;
    int workers = sms / 2;
    if (opt && opt->max_ctas > 0) workers = std::max(1, std::min(workers, opt->max_ctas / 2));
    workers = (int)std::max<int64_t>(1, std::min<int64_t>(workers, ceil_div((int64_t)P.ntri * ks, P.chunk_kt)));
    P.prof = nullptr;
    const bool timed = opt && opt->reserved[0] == 1;
    if (timed) {
      if (!g_ev_start) {
        TD_CHECK_CUDA(cudaEventCreate(&g_ev_start));
        TD_CHECK_CUDA(cudaEventCreate(&g_ev_stop));
      }
      TD_CHECK_CUDA(cudaEventRecord(g_ev_start, st));
    }
    auto kernel = f16 ? hk::lagged_gram_halo_kernel<true, true> : hk::lagged_gram_halo_kernel<true, false>;
    const size_t smem_bytes = f16 ? hk::F16_RECT_SMEM_BYTES : hk::RECT_SMEM_BYTES;
    TD_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * 2);
    cfg.blockDim = dim3(f16 ? hk::NUM_THREADS_F16 : hk::NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmap, tmap_a, P));
    TD_CHECK_CUDA(cudaGetLastError());
    if (timed) {
      TD_CHECK_CUDA(cudaEventRecord(g_ev_stop, st));
      g_ev_valid = true;
    }
    if ((rc = debug_sync(st, "toeplitz block-row kernel"))) return rc;
  }
  const int cblocks = (int)ceil_div(C, 32);
  const int nb = (int)ceil_div(pl.Dz, 32);
  // (batched: chunks of segments keep gridDim.x and gridDim.z inside their limits)
  for (int64_t o0 = 0; o0 < n_out; o0 += 16384) {
    const int64_t nz = std::min<int64_t>(16384, n_out - o0);
    dim3 grid((unsigned)(cblocks * nz), (unsigned)C, (unsigned)L);
    toeplitz_expand_kernel<<<grid, 32, 0, st>>>(Ct + o0 * pl.Dz * ldc, ldc, streams[0].data_dev, streams[0].ld, C, L, pre,
                                                reinterpret_cast<const SegEdge*>(ws + off_seg) + (batched ? o0 : 0), n_seg,
                                                G + (batched ? o0 * pl.Dz * ldg : 0), ldg, cblocks,
                                                batched ? pl.Dz * ldc : 0, batched ? pl.Dz * ldg : 0);
    TD_CHECK_CUDA(cudaGetLastError());
    symmetrize_kernel<<<dim3(nb, nb, (unsigned)nz), dim3(32, 8), 0, st>>>(G + (batched ? o0 * pl.Dz * ldg : 0), pl.Dz, ldg, 1,
                                                                          batched ? pl.Dz * ldg : 0);
    TD_CHECK_CUDA(cudaGetLastError());
  }
  return TD_OK;
}

}  // namespace
}  // namespace td

using namespace td;

extern "C" int td_lagged_gram_workspace_bytes(const td_stream* streams, int n_streams,
                                              const int64_t* seg_first, const int64_t* seg_keep,
                                              int n_seg, const td_gram_options* opt, size_t* bytes) {
  TD_REQUIRE(streams && seg_keep && bytes, "td_lagged_gram_workspace_bytes: null argument");
  if (opt && opt->method == 1) { *bytes = 0; return TD_OK; }
  if (opt && opt->method == 5)
    return run_toeplitz(streams, n_streams, seg_first, seg_keep, n_seg, nullptr, 0, opt, nullptr, 0, nullptr, bytes);
  bool halo = false;
  int rc = use_halo(streams, n_streams, opt, &halo);
  if (rc) return rc;
  if (halo) {
    HaloPlan hp;
    rc = make_halo_plan(streams, n_streams, seg_first, seg_keep, n_seg, &hp, halo_is_f16(opt));
    if (rc) return rc;
    *bytes = hp.total;
    return TD_OK;
  }
  GramPlan pl;
  rc = make_plan(streams, n_streams, seg_first, seg_keep, n_seg, opt, &pl);
  if (rc) return rc;
  *bytes = pl.total;
  return TD_OK;
}

namespace td {
void release_gram_events() {
  if (g_ev_start) cudaEventDestroy(g_ev_start);
  if (g_ev_stop) cudaEventDestroy(g_ev_stop);
  g_ev_start = g_ev_stop = nullptr;
  g_ev_valid = false;
}
}  // namespace td

extern "C" float td_last_gram_kernel_ms(void) {
  if (!g_ev_valid) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(g_ev_stop) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, g_ev_start, g_ev_stop) != cudaSuccess) return -1.0f;
  return ms;
}

extern "C" int td_symmetrize_f32(float* G, int64_t n, int64_t ldg, int from_upper, void* stream) {
  TD_REQUIRE(G && n >= 0 && ldg >= n, "td_symmetrize_f32: bad arguments");
  if (n == 0) return TD_OK;
  const int nb = (int)ceil_div(n, 32);
  symmetrize_kernel<<<dim3(nb, nb), dim3(32, 8), 0, (cudaStream_t)stream>>>(G, n, ldg, from_upper, 0);
  TD_CHECK_CUDA(cudaGetLastError());
  return TD_OK;
}

extern "C" int td_lagged_gram_f32(const td_stream* streams, int n_streams, const int64_t* seg_first,
                                  const int64_t* seg_keep, int n_seg, float* G, int64_t ldg,
                                  const td_gram_options* opt, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TD_REQUIRE(streams && seg_keep && G, "td_lagged_gram_f32: null argument");
  if (opt && opt->method == 1) {
    int64_t Dz = 0;
    for (int q = 0; q < n_streams; ++q) {
      int rc = check_stream(&streams[q], seg_keep, n_seg, "td_lagged_gram");
      if (rc) return rc;
      Dz += width_of(streams[q]);
    }
    TD_REQUIRE(ldg >= Dz, "td_lagged_gram_f32: ldg (%lld) < Dz (%lld)", (long long)ldg, (long long)Dz);
    return launch_gram_simt(streams, n_streams, seg_first, seg_keep, n_seg, G, ldg, st);
  }
  if (opt && opt->method == 5)
    return run_toeplitz(streams, n_streams, seg_first, seg_keep, n_seg, G, ldg, opt, workspace, workspace_bytes, st, nullptr);
  bool halo = false;
  int rc = use_halo(streams, n_streams, opt, &halo);
  if (rc) return rc;
  if (halo)
    return run_halo(streams, n_streams, seg_first, seg_keep, n_seg, G, ldg, opt, workspace, workspace_bytes, st);
  GramPlan pl;
  rc = make_plan(streams, n_streams, seg_first, seg_keep, n_seg, opt, &pl);
  if (rc) return rc;
  // Kernel flavour here: 2 = operands from shared memory, 3 = A operand from tensor memory (CTA pairs
  // only).  (Method 4 / the automatic choice for CTA pairs went to run_halo above.)
  int flavour = opt ? opt->method : 0;
  if (flavour == 0) flavour = (pl.cg == 2) ? 3 : 2;
  TD_REQUIRE(flavour == 2 || pl.cg == 2,
             "td_lagged_gram_f32: method 3 (A operand in tensor memory) needs cta_group 2");
  TD_REQUIRE(ldg >= pl.Dz, "td_lagged_gram_f32: ldg (%lld) < Dz (%lld)", (long long)ldg, (long long)pl.Dz);
  if (workspace_bytes < pl.total || !workspace) {
    set_error("td_lagged_gram_f32: workspace has %zu bytes, %zu needed", workspace_bytes, pl.total);
    return TD_ERR_WORKSPACE;
  }
  int dev = 0, sms = 0, cc_major = 0;
  TD_CHECK_CUDA(cudaGetDevice(&dev));
  TD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TD_CHECK_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_major != 10) {
    set_error("td_lagged_gram_f32: the tcgen05 kernel needs an sm_100 device (found sm_%d)", cc_major * 10);
    return TD_ERR_UNSUPPORTED;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  // ---- tables -------------------------------------------------------------------------------
  std::vector<int4> groups(pl.n_groups);
  std::vector<int> out_idx(pl.Dpad, -1);
  {
    int g = 0;
    int64_t col0 = 0;
    for (int q = 0; q < n_streams; ++q) {
      const int per_lag = pl.Cpad[q] / pl.CB;
      for (int k = 0; k < pl.L[q]; ++k)
        for (int b = 0; b < per_lag; ++b, ++g) {
          const int d = k - pl.pre[q];
          const int sh = (pl.n_shift == 4) ? ((d % 4) + 4) % 4 : 0;
          groups[g] = make_int4(q, d - sh, b * pl.CB, sh);
          for (int i = 0; i < pl.CB; ++i) {
            const int c = b * pl.CB + i;
            if (c < pl.C[q]) out_idx[(int64_t)g * pl.CB + i] = (int)(col0 + (int64_t)k * pl.C[q] + c);
          }
        }
      col0 += width_of(streams[q]);
    }
    for (; g < pl.n_groups; ++g) groups[g] = make_int4(0, 0, pl.Cpad[0], 0);   // fully out of bounds -> zeros
  }
  std::vector<int2> ktiles;
  ktiles.reserve(pl.NK);
  for (int s = 0; s < n_seg; ++s) {
    const int64_t f = first_of(seg_first, s);
    for (int64_t j = f / BK; j * BK < seg_keep[s] && seg_keep[s] > f; ++j) {
      const int begin = (int)std::max<int64_t>(0, f - j * BK);
      const int end = (int)std::min<int64_t>(BK, seg_keep[s] - j * BK);
      ktiles.push_back(make_int2((int)(pl.seg_off[s] + j * BK), begin | (end << 8)));
    }
  }
  std::vector<int64_t> segtab((size_t)n_seg * 2 * n_streams);
  for (int q = 0; q < n_streams; ++q)
    for (int s = 0; s < n_seg; ++s) {
      segtab[((size_t)q * 2 + 0) * n_seg + s] = streams[q].seg_start[s];
      segtab[((size_t)q * 2 + 1) * n_seg + s] = streams[q].seg_len[s];
    }
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_groups, groups.data(), groups.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
  if (!ktiles.empty())
    TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_ktiles, ktiles.data(), ktiles.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_outidx, out_idx.data(), out_idx.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_segoff, pl.seg_off.data(), (size_t)n_seg * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  TD_CHECK_CUDA(cudaMemcpyAsync(ws + pl.off_segtab, segtab.data(), segtab.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));

  // ---- stage ----------------------------------------------------------------------------------
  CUtensorMap tmaps[2];
  for (int q = 0; q < n_streams; ++q) {
    float* planes = reinterpret_cast<float*>(ws + pl.off_stage[q]);
    TD_CHECK_CUDA(cudaMemsetAsync(planes, 0, (size_t)pl.n_shift * pl.Cpad[q] * pl.Tp * sizeof(float), st));
    StreamView sv;
    sv.data = streams[q].data_dev;
    sv.ld = streams[q].ld;
    sv.channels = streams[q].channels;
    sv.pre = streams[q].pre;
    sv.lags = pl.L[q];
    const int64_t* tab = reinterpret_cast<const int64_t*>(ws + pl.off_segtab);
    sv.seg_start = tab + ((size_t)q * 2 + 0) * n_seg;
    sv.seg_len = tab + ((size_t)q * 2 + 1) * n_seg;
    int64_t maxlen = 0;
    for (int s = 0; s < n_seg; ++s) maxlen = std::max(maxlen, streams[q].seg_len[s]);
    if (maxlen > 0) {
      for (int s0 = 0; s0 < n_seg; s0 += 32768) {
        dim3 grid((unsigned)ceil_div(maxlen, 32), (unsigned)std::min(32768, n_seg - s0),
                  (unsigned)ceil_div(pl.C[q], 32));
        stage_kernel<<<grid, dim3(32, 8), 0, st>>>(
            sv, reinterpret_cast<const int64_t*>(ws + pl.off_segoff), s0, planes, pl.Tp, pl.Cpad[q], pl.n_shift);
        TD_CHECK_CUDA(cudaGetLastError());
      }
    }
    rc = make_tmap(&tmaps[q], planes, pl.Tp, pl.Cpad[q], pl.CB, pl.n_shift);
    if (rc) return rc;
  }
  if ((rc = debug_sync(st, "stage"))) return rc;
  if (n_streams == 1) tmaps[1] = tmaps[0];

  // ---- gram -----------------------------------------------------------------------------------
  TD_CHECK_CUDA(cudaMemset2DAsync(G, (size_t)ldg * sizeof(float), 0, (size_t)pl.Dz * sizeof(float), (size_t)pl.Dz, st));
  if (pl.NK > 0) {
    GramParams P;
    P.groups = reinterpret_cast<const int4*>(ws + pl.off_groups);
    P.ktiles = reinterpret_cast<const int2*>(ws + pl.off_ktiles);
    P.out_idx = reinterpret_cast<const int*>(ws + pl.off_outidx);
    P.G = G;
    P.ldg = ldg;
    P.CB = pl.CB;
    P.groups_per_tile = TILE_ROWS / pl.CB;
    P.ntri = pl.ntri;
    P.NK = pl.NK;
    int chunk_rows = (opt && opt->chunk_rows > 0) ? opt->chunk_rows : 128;
    P.chunk_kt = std::max(1, chunk_rows / BK);
    // Superchunk: the staged frames all tiles sweep together; sized to stay L2 resident (~24 MB).
    int64_t sc_rows = (opt && opt->superchunk_rows > 0) ? opt->superchunk_rows : 0;
    if (sc_rows == 0) {
      int64_t cpad_sum = 0;
      for (int q = 0; q < n_streams; ++q) cpad_sum += pl.Cpad[q];
      sc_rows = (24LL << 20) / (cpad_sum * 4 * pl.n_shift);
      sc_rows = std::max<int64_t>(sc_rows, 4096);
    }
    int64_t ks = std::max<int64_t>(P.chunk_kt, round_up(ceil_div(sc_rows, BK), P.chunk_kt));
    ks = std::min<int64_t>(ks, round_up(pl.NK, P.chunk_kt));
    P.KS = (int)ks;
    const int64_t nsc = ceil_div(pl.NK, ks);
    P.total_steps = nsc * pl.ntri * ks;
    int workers = sms / pl.cg;
    if (opt && opt->max_ctas > 0) workers = std::max(1, std::min(workers, opt->max_ctas / pl.cg));
    // No more workers than there are chunk-sized pieces of work in a superchunk.
    workers = (int)std::max<int64_t>(1, std::min<int64_t>(workers, ceil_div((int64_t)pl.ntri * ks, P.chunk_kt)));
    P.prof = nullptr;
#ifdef TD_TS_PROFILE
    long long* prof_dev = nullptr;
    TD_CHECK_CUDA(cudaMalloc(&prof_dev, (size_t)sms * 24 * sizeof(long long)));
    TD_CHECK_CUDA(cudaMemsetAsync(prof_dev, 0, (size_t)sms * 24 * sizeof(long long), st));
    P.prof = prof_dev;
#endif
    const bool timed = opt && opt->reserved[0] == 1;
    if (timed) {
      if (!g_ev_start) {
        TD_CHECK_CUDA(cudaEventCreate(&g_ev_start));
        TD_CHECK_CUDA(cudaEventCreate(&g_ev_stop));
      }
      TD_CHECK_CUDA(cudaEventRecord(g_ev_start, st));
    }
    if (flavour == 3) {
      auto kernel = ts::lagged_gram_ts_kernel<2>;          // two MMA issuing warps
      TD_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts::SMEM_BYTES));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(workers * 2);
      cfg.blockDim = dim3(ts::NUM_THREADS);
      cfg.dynamicSmemBytes = ts::SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      TD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmaps[0], tmaps[1], P));
    } else if (pl.cg == 1) {
      TD_CHECK_CUDA(cudaFuncSetAttribute(lagged_gram_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      lagged_gram_umma_kernel<1><<<workers, NUM_THREADS, SMEM_BYTES, st>>>(tmaps[0], tmaps[1], P);
    } else {
      TD_CHECK_CUDA(cudaFuncSetAttribute(lagged_gram_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(workers * 2);
      cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      TD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lagged_gram_umma_kernel<2>, tmaps[0], tmaps[1], P));
    }
    TD_CHECK_CUDA(cudaGetLastError());
    if (timed) {
      TD_CHECK_CUDA(cudaEventRecord(g_ev_stop, st));
      g_ev_valid = true;
    }
#ifdef TD_TS_PROFILE
    {
      std::vector<long long> h((size_t)sms * 24);
      TD_CHECK_CUDA(cudaStreamSynchronize(st));
      TD_CHECK_CUDA(cudaMemcpy(h.data(), prof_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      cudaFree(prof_dev);
      const char* names[24] = {"prod wait empty", "mma1 wait ready", "mma0 wait ready", "mma0 wait tempty", "mma1 wait tempty",
                               "mma1 other", "mma0 other", "mma1 issue+commit", "conv wait full", "conv lds+cvt+sts", "conv wait aempty",
                               "conv loop top", "conv st+arrive", "mma1 total", "mma0 total", "prod total", "mma0 issue+commit",
                               "conv total", "epi wait tfull", "epi drain", "epi store+other", "epi total", "", ""};
      for (int r = 0; r < 2; ++r) {
        fprintf(stderr, "[tdgram prof] rank %d (avg Mclk over CTAs):", r);
        for (int i = 0; i < 24; ++i) {
          if (!names[i][0]) continue;
          double sum = 0; int n = 0;
          for (int b = r; b < workers * 2; b += 2) { sum += (double)h[(size_t)b * 24 + i]; ++n; }
          fprintf(stderr, "  %s=%.2f", names[i], sum / n / 1e6);
        }
        fprintf(stderr, "\n");
      }
    }
#endif
    if (debug_enabled())
      fprintf(stderr, "[tdgram] gram launch: flavour=%d cg=%d workers=%d CB=%d Dz=%lld Dpad=%lld ntri=%d NK=%lld KS=%d chunk_kt=%d total=%lld Tp=%lld\n",
              flavour, pl.cg, workers, pl.CB, (long long)pl.Dz, (long long)pl.Dpad, pl.ntri, (long long)pl.NK, P.KS, P.chunk_kt,
              (long long)P.total_steps, (long long)pl.Tp);
    if ((rc = debug_sync(st, "gram kernel"))) return rc;
  }
  return td_symmetrize_f32(G, pl.Dz, ldg, /*from_upper=*/1, stream);
}

extern "C" int td_lagged_gram_batched_f32(const td_stream* streams, int n_streams, const int64_t* seg_first,
                                          const int64_t* seg_keep, int n_seg, float* G, int64_t ldg,
                                          const td_gram_options* opt, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  TD_REQUIRE(streams && seg_keep && G, "td_lagged_gram_batched_f32: null argument");
  td_gram_options o = opt ? *opt : td_gram_options{};
  TD_REQUIRE(o.method == 0 || o.method == 4 || o.method == 5 || o.method == 6,
             "td_lagged_gram_batched_f32: methods 4 (0 = default), 5 and 6 write per-segment Grams");
  if (o.method == 5)
    return run_toeplitz(streams, n_streams, seg_first, seg_keep, n_seg, G, ldg, &o, workspace, workspace_bytes,
                        (cudaStream_t)stream, nullptr, /*batched=*/true);
  if (o.method != 4) o.method = 6;
  o.cta_group = 2;
  return run_halo(streams, n_streams, seg_first, seg_keep, n_seg, G, ldg, &o, workspace, workspace_bytes,
                  (cudaStream_t)stream, /*batched=*/true);
}

extern "C" int td_lagged_gram_batched_workspace_bytes(const td_stream* streams, int n_streams,
                                                      const int64_t* seg_first, const int64_t* seg_keep,
                                                      int n_seg, const td_gram_options* opt, size_t* bytes) {
  TD_REQUIRE(streams && seg_keep && bytes, "td_lagged_gram_batched_workspace_bytes: null argument");
  if (opt && opt->method == 5)
    return run_toeplitz(streams, n_streams, seg_first, seg_keep, n_seg, nullptr, 0, opt, nullptr, 0, nullptr, bytes,
                        /*batched=*/true);
  td_gram_options o = opt ? *opt : td_gram_options{};
  if (o.method != 4) o.method = 6;
  o.cta_group = 2;
  return td_lagged_gram_workspace_bytes(streams, n_streams, seg_first, seg_keep, n_seg, &o, bytes);
}
