// CTA-pair Gram kernel with the A operand in tensor memory (included by gram_umma.cu).
//
// The shared-memory kernel (lagged_gram_umma_kernel) is bound by shared-memory bandwidth: per k-tile
// its 12 MMAs read ~144 KB of operands, the TMA writes 32 KB and the converters move 64 KB.  Here
// the A operand never returns to shared memory: four converter warps (one per TMEM lane quadrant)
// read each raw A row once, and write hi (the raw bits, kind::tf32 truncates) and lo straight into
// TMEM with tcgen05.st; tcgen05.mma takes A from TMEM and only B (raw + lo tile) from shared memory.
//
// TMEM (512 columns): three 128-column accumulator regions + a ring of four 32-column A slots
// (16 frames each: 16 hi + 16 lo columns).  The 256x256 pair tile is two 128-column halves; the halves
// of successive chunks rotate through the three regions (chunk c: half 0 -> region 2c mod 3, half 1
// -> region 2c+1 mod 3): half 0 of the next chunk always finds a region drained long ago and half 1
// gets the region half 0 has just left, so once half 1 trails half 0 by one drain time the drain
// into fp32 registers (every `chunk` k-tiles, bounds the tensor-core accumulation length) never
// stalls the tensor pipe although the accumulator is not double buffered.
//
// One thread cannot issue the MMAs fast enough (N = 128, K = 8: 64 tensor clocks per instruction,
// measured ~30 clocks of issue work each plus waits and commits), so each half has its own issuing
// warp; the halves write different regions and need no ordering between them.
//
//   warp 0      TMA producer (raw A tile, raw B tile; 4-stage ring of 48 KB)
//   warp 1, 2   MMA issuers for half 0 / half 1 (pair leader only): per k-step hi*hi + hi*lo + lo*hi
//   warps 4-7   converters: raw A rows -> TMEM (hi | lo, frames outside [begin, end) zeroed) and
//               raw B tile -> lo tile in shared memory, 16 frames at a time
//   warps 8-15  epilogue: TMEM region -> fp32 registers, red.add into G at the end of a work item
// Converters and epilogue warps of both CTAs arrive directly on the leader's barriers.

namespace ts {

// -DTD_TS_PROFILE: per-role stall clocks written to P.prof[blockIdx.x * 24 + i] (development aid).
#ifdef TD_TS_PROFILE
#define TS_LAP(i) { const long long n__ = clock64(); prof_acc[i] += n__ - lap__; lap__ = n__; }
#define TS_PROF_DECL long long lap__ = clock64(); long long prof_acc[24] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long prof_t0 = lap__;
#define TS_PROF_FLUSH(total_idx, ...) { if (lane == 0 && P.prof) { prof_acc[total_idx] = clock64() - prof_t0; const int idx__[] = {__VA_ARGS__, total_idx}; for (int i__ : idx__) P.prof[blockIdx.x * 24 + i__] = prof_acc[i__]; } }
#else
#define TS_LAP(i)
#define TS_PROF_DECL
#define TS_PROF_FLUSH(total_idx, ...) {}
#endif

constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 3 * OPERAND_BYTES;           // A_raw, B_raw (hi), B_lo
constexpr int NUM_THREADS = 512;
constexpr int NUM_CONV = 4;                              // warps 4..7
constexpr int EPI_WARP0 = 8, NUM_EPI = 8;
constexpr uint32_t ACC_COLS = 128;                       // columns per accumulator region
constexpr uint32_t A_COL0 = 3 * ACC_COLS;                // first column of the A ring
constexpr uint32_t A_SLOT_COLS = 32;                     // 16 frames: 16 hi + 16 lo columns
constexpr int A_SLOTS = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr int NACC = 128;                                // accumulator values per epilogue thread
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_BARRIERS = 2 * STAGES + 2 * A_SLOTS + 12;
static_assert(8 * NUM_BARRIERS + 8 <= 256, "barrier block too small");

// Region bookkeeping of one half (h = 0, 1) of the pair tile, kept identically by its MMA issuer and
// its epilogue warps: chunk number cm (mod 3) picks the region, `own` bit r is the parity of this
// half's next use of region r.  Every mbarrier of the accumulator hand-off belongs to ONE (half,
// region) pair, so each has one waiting role that waits its phases strictly in order -- with one
// barrier per region shared by both halves a half that runs two short chunks ahead would wait for a
// phase whose predecessor has not completed yet, which the parity test reports as "done".
struct HalfRing {
  int cm = 0;
  uint32_t own = 0;
  int r;
  uint32_t n;
  __device__ __forceinline__ void next_chunk(int h) {
    r = (2 * cm + h) % 3;
    n = (own >> r) & 1u;
    own ^= 1u << r;
    cm = (cm + 1 == 3) ? 0 : cm + 1;
  }
  // Parity to wait for on the OTHER half's "drained" barrier of region r before this use: region 1 is
  // first used by half 1, regions 0 and 2 by half 0; the first user waits for the other half's previous
  // drain (one phase back), the second user for the drain of the same ordinal.
  __device__ __forceinline__ uint32_t drained_parity(int h) const {
    const bool first_user = (r == 1) ? (h == 1) : (h == 0);
    return first_user ? (n ^ 1u) : n;
  }
};

// lo = tf32_rna(x - trunc_tf32(x)); the subtraction is exact in fp32.  Round-to-nearest = add half a
// tf32 ulp to the magnitude bits; the 13 low bits are left as they fall, kind::tf32 does not read them.
__device__ __forceinline__ uint32_t lo_bits(uint32_t x) {
  const float d = __uint_as_float(x) - __uint_as_float(x & 0xFFFFE000u);
  return __float_as_uint(d) + 0x1000u;
}

template <int ISSUERS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
lagged_gram_ts_kernel(const __grid_constant__ CUtensorMap tmap0,
                      const __grid_constant__ CUtensorMap tmap1, const GramParams P) {
  constexpr uint32_t IDESC = ptx::umma_idesc_tf32(256, 128);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // raw tiles landed (local)
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };           // MMAs of the stage retired (both issuers)
  auto ready_bar = [&](int j) { return bar_base + 8u * (2 * STAGES + j); };       // leader: A slot + B lo of 16 frames built in both CTAs
  auto aempty_bar = [&](int j) { return bar_base + 8u * (2 * STAGES + A_SLOTS + j); };   // MMAs reading the slot retired
  auto tfull_bar = [&](int h, int r) { return bar_base + 8u * (2 * STAGES + 2 * A_SLOTS + 3 * h + r); };        // half h finished a chunk in region r
  auto tempty_bar = [&](int h, int r) { return bar_base + 8u * (2 * STAGES + 2 * A_SLOTS + 6 + 3 * h + r); };   // leader: half h's chunk drained from region r
  const uint32_t tmem_slot = bar_base + 8u * NUM_BARRIERS;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int worker = blockIdx.x / 2;
  const int n_workers = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap0);
    ptx::prefetch_tmap(&tmap1);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), ISSUERS);
    }
    for (int j = 0; j < A_SLOTS; ++j) {
      ptx::mbar_init(ready_bar(j), 2 * NUM_CONV);
      ptx::mbar_init(aempty_bar(j), ISSUERS);
    }
    for (int h = 0; h < 2; ++h)
      for (int r = 0; r < 3; ++r) {
        ptx::mbar_init(tfull_bar(h, r), 1);
        ptx::mbar_init(tempty_bar(h, r), 2 * (NUM_EPI / 2));
      }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<2>(tmem_slot, TMEM_COLS);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Register budget per warpgroup: 128 x (72 + 104 + 168 + 168) = 65536.
  const int wg = warp >> 2;
  if (wg == 0) {
    ptx::setmaxnreg_dec<72>();
    if (warp == 0) {
      // ===================== TMA producer (whole warp loops, one elected lane issues) ============
      WorkIter wi(P, worker, n_workers);
      WorkItem it;
      int stage = 0;
      uint32_t phase = 0;
      TS_PROF_DECL
      while (wi.next(P, it)) {
        const int ga = (it.tm * 2 + (int)rank) * P.groups_per_tile;
        const int gb = (it.tn * 2 + (int)rank) * P.groups_per_tile;
        int kt_time = __ldg(&P.ktiles[it.kt_begin]).x;
        for (long long kt = it.kt_begin; kt < it.kt_end; ++kt) {
          const int t0 = kt_time;
          if (kt + 1 < it.kt_end) kt_time = __ldg(&P.ktiles[kt + 1]).x;
          TS_LAP(1)
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          TS_LAP(0)
          if (ptx::elect_one()) {
            const uint32_t fb = full_bar(stage);
            ptx::mbar_expect_tx(fb, 2 * OPERAND_BYTES);
            const uint32_t sbase = smem_base + stage * STAGE_BYTES;
#pragma unroll 1
            for (int op = 0; op < 2; ++op) {
              const int g0 = op ? gb : ga;
              const uint32_t obase = sbase + op * OPERAND_BYTES;
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
      TS_PROF_FLUSH(15, 0)
    } else if (warp <= ISSUERS && rank == 0) {
      // ===================== MMA issuers (pair leader) ==========================================
      // ISSUERS == 2: warp 1 -> half 0, warp 2 -> half 1.  ISSUERS == 1: warp 1 issues both halves.
      const int h_begin = (ISSUERS == 2) ? warp - 1 : 0;
      const int h_end = (ISSUERS == 2) ? warp : 2;
      WorkIter wi(P, worker, n_workers);
      WorkItem it;
      int stage = 0;
      uint32_t kcount = 0;                       // k-tiles issued so far: A slots 2*kcount, 2*kcount + 1
      HalfRing ring_a, ring_b;                   // ISSUERS == 2: ring_a is this warp's half; 1: halves 0 and 1
      TS_PROF_DECL
      while (wi.next(P, it)) {
        int next_range = __ldg(&P.ktiles[it.kt_begin]).y;
        for (long long c0 = it.kt_begin; c0 < it.kt_end; c0 += P.chunk_kt) {
          const long long c1 = min(it.kt_end, c0 + (long long)P.chunk_kt);
          ring_a.next_chunk(h_begin);
          if (ISSUERS == 1) ring_b.next_chunk(1);
          for (long long kt = c0; kt < c1; ++kt, ++kcount) {
            const int range = next_range;
            if (kt + 1 < it.kt_end) next_range = __ldg(&P.ktiles[kt + 1]).y;
            const int ks0 = (range & 0xff) >> 3, ks1 = ((range >> 8) + 7) >> 3;
            const bool whole = (ks0 == 0) && (ks1 == BK / 8);
            const uint32_t sbase = smem_base + stage * STAGE_BYTES;
            const uint32_t first_acc = (kt > c0) ? 1u : 0u;
            for (int h = h_begin; h < h_end; ++h) {
              const HalfRing& hr = (ISSUERS == 2 || h == 0) ? ring_a : ring_b;
              const int r = hr.r;
              const uint32_t d = tmem_base + (uint32_t)r * ACC_COLS;
              if (kt == c0) {
                TS_LAP(6)
                ptx::mbar_wait(tempty_bar(1 - h, r), hr.drained_parity(h));
                TS_LAP(3)
              }
              const uint64_t b_hi = ptx::umma_desc_k_sw128(sbase + OPERAND_BYTES + h * (OPERAND_BYTES / 2));
              const uint64_t b_lo = ptx::umma_desc_k_sw128(sbase + 2 * OPERAND_BYTES + h * (OPERAND_BYTES / 2));
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint32_t use = 2u * kcount + hh;
                const uint32_t slot = use & (A_SLOTS - 1);
                if (ISSUERS == 2 || h == 0) {
                  TS_LAP(6)
                  ptx::mbar_wait(ready_bar(slot), (use >> 2) & 1u);
                  TS_LAP(2)
                }
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                  const uint32_t a_base = tmem_base + A_COL0 + slot * A_SLOT_COLS;
                  if (whole) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                      const uint64_t adv = (uint64_t)((hh * 2 + q) * 2);     // 8 frames = 32 B = 2 x 16 B
                      const uint32_t a_hi = a_base + q * 8, a_lo = a_base + 16 + q * 8;
                      ptx::mma_tf32_ts_pair(d, a_hi, b_hi + adv, IDESC, (hh + q) ? 1u : first_acc);
                      ptx::mma_tf32_ts_pair(d, a_hi, b_lo + adv, IDESC, 1u);
                      ptx::mma_tf32_ts_pair(d, a_lo, b_hi + adv, IDESC, 1u);
                    }
                  } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                      const int ks = hh * 2 + q;
                      if (ks >= ks0 && ks < ks1) {
                        const uint64_t adv = (uint64_t)(ks * 2);
                        const uint32_t a_hi = a_base + q * 8, a_lo = a_base + 16 + q * 8;
                        ptx::mma_tf32_ts_pair(d, a_hi, b_hi + adv, IDESC, (ks > ks0) ? 1u : first_acc);
                        ptx::mma_tf32_ts_pair(d, a_hi, b_lo + adv, IDESC, 1u);
                        ptx::mma_tf32_ts_pair(d, a_lo, b_hi + adv, IDESC, 1u);
                      }
                    }
                  }
                  if (ISSUERS == 2 || h == 1) ptx::mma_commit<2>(aempty_bar(slot));   // A slot reusable
                  if (hh == 1) {
                    if (ISSUERS == 2 || h == 1) ptx::mma_commit<2>(empty_bar(stage)); // stage reusable in both CTAs
                    if (kt == c1 - 1) ptx::mma_commit<2>(tfull_bar(h, r));
                  }
                }
                __syncwarp();
                TS_LAP(16)
              }
            }
            if (++stage == STAGES) stage = 0;
          }
        }
      }
#ifdef TD_TS_PROFILE
      if (lane == 0 && P.prof) {
        long long* o = P.prof + blockIdx.x * 24;
        const long long total = clock64() - prof_t0;
        if (warp == 1) { o[2] = prof_acc[2]; o[3] = prof_acc[3]; o[6] = prof_acc[6]; o[16] = prof_acc[16]; o[14] = total; }
        else { o[1] = prof_acc[2]; o[4] = prof_acc[3]; o[5] = prof_acc[6]; o[7] = prof_acc[16]; o[13] = total; }
      }
#endif
    }
  } else if (wg == 1) {
    // ===================== converters: A rows -> TMEM (hi | lo), B tile -> lo tile ==============
    ptx::setmaxnreg_dec<104>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                    // A row (= TMEM lane) and B row of this thread
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t ready_leader0 = ptx::mapa(ready_bar(0), 0);
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + A_COL0;
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t kcount = 0;
    TS_PROF_DECL
    while (wi.next(P, it)) {
      int next_range = __ldg(&P.ktiles[it.kt_begin]).y;
      for (long long kt = it.kt_begin; kt < it.kt_end; ++kt, ++kcount) {
        const int range = next_range;
        if (kt + 1 < it.kt_end) next_range = __ldg(&P.ktiles[kt + 1]).y;
        const int begin = range & 0xff, end = range >> 8;
        const uint32_t araw = smem_base + stage * STAGE_BYTES + row_off;
        const uint32_t braw = araw + OPERAND_BYTES, blo = araw + 2 * OPERAND_BYTES;
        TS_LAP(11)
        ptx::mbar_wait(full_bar(stage), phase);
        TS_LAP(8)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t use = 2u * kcount + hh;
          const uint32_t slot = use & (A_SLOTS - 1);
          // Frame t of a row sits in 16-byte chunk (t/4) ^ (row%8) (128-byte swizzle): lane-per-row
          // 16-byte accesses are conflict free.
          uint32_t hi[16], lo[16], bv[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t off = ((uint32_t)(hh * 4 + i) ^ sw) << 4;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(hi[4 * i]), "=r"(hi[4 * i + 1]), "=r"(hi[4 * i + 2]), "=r"(hi[4 * i + 3])
                         : "r"(araw + off));
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(bv[4 * i]), "=r"(bv[4 * i + 1]), "=r"(bv[4 * i + 2]), "=r"(bv[4 * i + 3])
                         : "r"(braw + off));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t off = ((uint32_t)(hh * 4 + i) ^ sw) << 4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blo + off),
                         "r"(lo_bits(bv[4 * i])), "r"(lo_bits(bv[4 * i + 1])), "r"(lo_bits(bv[4 * i + 2])),
                         "r"(lo_bits(bv[4 * i + 3]))
                         : "memory");
          }
          ptx::fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor-core proxy
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int t = hh * 16 + e;
            if (t < begin || t >= end) hi[e] = 0u;
            lo[e] = lo_bits(hi[e]);
          }
          TS_LAP(9)
          ptx::mbar_wait(aempty_bar(slot), ((use >> 2) & 1u) ^ 1u);
          TS_LAP(10)
          ptx::tc_fence_after();
          ptx::tmem_st16(tlane + slot * A_SLOT_COLS, hi);
          ptx::tmem_st16(tlane + slot * A_SLOT_COLS + 16, lo);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(ready_leader0 + 8u * slot);
          TS_LAP(12)
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (warp == 4) TS_PROF_FLUSH(17, 8, 9, 10, 11, 12)
  } else {
    // ===================== epilogue: TMEM region -> fp32 registers -> red.add =====================
    ptx::setmaxnreg_inc<168>();
    const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - EPI_WARP0) >> 2;  // which 128-column half of the pair tile
    const uint32_t tempty_leader0 = ptx::mapa(tempty_bar(half, 0), 0);
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    HalfRing ring;
    float acc[NACC];
    TS_PROF_DECL
    while (wi.next(P, it)) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
      for (long long c0 = it.kt_begin; c0 < it.kt_end; c0 += P.chunk_kt) {
        ring.next_chunk(half);
        const int r = ring.r;
        TS_LAP(20)
        ptx::mbar_wait(tfull_bar(half, r), ring.n);
        TS_LAP(18)
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)r * ACC_COLS;
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
        if (lane == 0) ptx::mbar_arrive_remote(tempty_leader0 + 8u * r);
        TS_LAP(19)
      }
      // Tile (tm, tn), tm >= tn: element (m, n) is stored at G[n][m] (upper triangle) so that the
      // 32 lanes of a warp (consecutive m) write consecutive addresses.  Accumulator column i of half
      // h is B row h*64 + i%64 of CTA i/64 (the pair's N dimension is split between the CTAs).
      const int m_row = (it.tm * 2 + (int)rank) * TILE_ROWS + quad * 32 + lane;
      const int m_out = __ldg(&P.out_idx[m_row]);
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        const int n_row = it.tn * 256 + (i >> 6) * TILE_ROWS + half * 64 + (i & 63);
        const int n_out = __ldg(&P.out_idx[n_row]);
        if (m_out >= 0 && n_out >= 0 && n_row <= m_row)
          atomicAdd(P.G + (long long)n_out * P.ldg + m_out, acc[i]);
      }
      TS_LAP(20)
    }
    if (warp == EPI_WARP0) TS_PROF_FLUSH(21, 18, 19, 20)
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc<2>(tmem_base, TMEM_COLS);
}

}  // namespace ts
