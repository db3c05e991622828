// CTA-pair Gram kernel that builds the lagged operand ON CHIP from a raw time tile + lag halo
// (included by gram_umma.cu).  Two flavours of one template: 3xTF32 on the caller's fp32 array
// (method 4, described first) and fp16 x 3 on a pre-split copy (method 6, the default for CTA pairs;
// see the template's comment and DESIGN.md section 3 for what differs).
//
// The earlier kernels (gram_ts.cuh, lagged_gram_umma_kernel) fetch every (lag, channel-block) row
// group of an operand tile as its own TMA box from four shifted channel-major copies of the
// recording: 4x the recording in HBM, and each raw sample crosses L2 -> shared memory once per lag
// and tile.  Here the recording stays as the caller has it -- time-major [frames, channels] -- and
//
//   * an operand row group is 32 rows = 4 consecutive lags x 8 channels of one stream.  What it needs
//     for a 32-frame k-tile is ONE raw box [32 + halo frames, 8 channels] straight from the raw
//     array (TMA, no swizzle, 32 bytes per frame); consecutive lag groups of the same channels share
//     the box of the first one (16 lags -> 47 frames, boxes are 48 frames), so a 128-row operand
//     tile needs 1.5 - 6 KB of raw samples instead of 16 KB;
//   * the four converter warps (one per TMEM lane quadrant, lane = 4 lags x 8 channels) read their
//     row's 16 frames at the row's own lag offset -- consecutive lanes hit 32 consecutive words, no
//     bank conflicts -- zero what lies outside the trial (brain_data.py:448-452 pads with zeros at
//     trial / chunk edges; a TMA box simply runs into the neighbouring trial) or outside this call's
//     design rows, and write
//         A: hi | lo straight into TENSOR memory (tcgen05.st), as in gram_ts.cuh,
//         B: the hi tile and the lo tile in the K-major 128-byte-swizzled layout tcgen05.mma reads;
//   * MMA issue, accumulator regions, epilogue and barriers are those of gram_ts.cuh.
//
// The Hankel matrix never exists anywhere: HBM holds the raw recording once (a padded copy only if
// the caller's row stride is not a multiple of 16 bytes, which TMA requires; the fp16 x 3 flavour
// reads its pre-split copy instead -- the same 4 bytes per sample).

namespace hk {

#ifndef TD_HALO_STAGES
#define TD_HALO_STAGES 4
#endif
constexpr int STAGES = TD_HALO_STAGES;
constexpr int BOX_FRAMES = 48;                          // 32 frames + 15 frames of lag halo + 1
constexpr int BOX_CH = 8;                               // channels per row group
constexpr int BOX_BYTES = BOX_FRAMES * BOX_CH * 4;      // 1536
constexpr int RG_ROWS = 32;                             // rows per row group: 4 lags x 8 channels
constexpr int RG_LAGS = 4;
constexpr int RG_PER_CTA = TILE_ROWS / RG_ROWS;         // 4
constexpr int RAW_BYTES = RG_PER_CTA * BOX_BYTES;       // raw boxes of one operand tile
constexpr int OFF_BHI = 0;
constexpr int OFF_BLO = OPERAND_BYTES;
constexpr int OFF_ARAW = 2 * OPERAND_BYTES;
constexpr int OFF_BRAW = OFF_ARAW + RAW_BYTES;
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES + 2 * RAW_BYTES;     // 44 KiB
static_assert(STAGE_BYTES % 1024 == 0, "stage must keep the 1024-byte alignment of the swizzled tiles");
// Rectangular flavour (method 5: A = the lag-0 rows only): an A row group is 1 lag x 32 channels, its raw
// box 32 frames x 32 channels (no halo: one lag), so all 128 A rows of a CTA are real channels.
constexpr int ABOX_CH = 32;
constexpr int ABOX_FRAMES = 32;
constexpr int ABOX_BYTES = ABOX_FRAMES * ABOX_CH * 4;    // 4096
constexpr int RECT_RAW_A_BYTES = RG_PER_CTA * ABOX_BYTES;                      // 16 KiB
constexpr int RECT_OFF_BRAW = OFF_ARAW + RECT_RAW_A_BYTES;
constexpr int RECT_STAGE_BYTES = 2 * OPERAND_BYTES + RECT_RAW_A_BYTES + RAW_BYTES;   // 54 KiB
static_assert(RECT_STAGE_BYTES % 1024 == 0, "stage must keep the 1024-byte alignment of the swizzled tiles");
constexpr size_t RECT_SMEM_BYTES = (size_t)STAGES * RECT_STAGE_BYTES + 1024 + 256;
static_assert(RECT_SMEM_BYTES <= 232448, "rectangular flavour exceeds the 227 KB of shared memory");
constexpr int NUM_THREADS = 512;
constexpr int NUM_THREADS_F16 = 640;                     // fp16 x 3: one more warpgroup (B converters)
constexpr int NUM_CONV = 4;                              // warps 4..7 (fp16 x 3: 4..7 build A, 8..11 build B)
constexpr int NUM_EPI = 8;                               // the last two warpgroups
constexpr uint32_t ACC_COLS = 128;
constexpr uint32_t A_COL0 = 3 * ACC_COLS;
constexpr uint32_t A_SLOT_COLS = 32;
constexpr int A_SLOTS = 4;
constexpr uint32_t TMEM_COLS = 512;
constexpr int NACC = 128;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_BARRIERS = 2 * STAGES + 2 * A_SLOTS + 12;
static_assert(8 * NUM_BARRIERS + 8 <= 256, "barrier block too small");
// fp16 x 3 flavour: a B row holds its 32 hi halves in bytes [0, 64) and its 32 lo halves in bytes
// [64, 128) of ONE 128-byte swizzled row, so a stage is 16 KiB of B + 12 KiB of raw boxes and eight
// stages fit.  The raw ring has to be deep: a stage is busy from the TMA issue to the retirement of the
// MMAs that read it (TMA latency + conversion + MMA + commit, ~4000 clocks), a k-tile lasts 768.
// Both flavours run at the board's power limit, so a warp that only waits should not spin: the TMA
// producer (its ring is normally full), the epilogue warps (idle for most of a chunk) and the MMA
// issuers (they wait for the converters 9 polls out of 10) sleep between polls.  Measured at the cfg 4
// width: 703 -> 686 ms with the first two, -> 667 ms with the issuers (any value from 20 to 200 ns);
// sleeping converter waits change nothing (the converters rarely wait).
#ifndef TD_F16_SLEEP_NS
#define TD_F16_SLEEP_NS 100
#endif
#ifndef TD_F16_ISSUER_SLEEP_NS
#define TD_F16_ISSUER_SLEEP_NS 40
#endif
#ifndef TD_F16_STAGES
#define TD_F16_STAGES 6
#endif
constexpr int F16_STAGES = TD_F16_STAGES;
constexpr int F16_OFF_BLO = 64;
constexpr int F16_OFF_ARAW = OPERAND_BYTES;
constexpr int F16_OFF_BRAW = F16_OFF_ARAW + RAW_BYTES;
constexpr int F16_STAGE_BYTES = OPERAND_BYTES + 2 * RAW_BYTES;      // 28 KiB
static_assert(F16_STAGE_BYTES % 1024 == 0, "stage must keep the 1024-byte alignment of the swizzled tiles");
constexpr int F16_NUM_BARRIERS = 2 * F16_STAGES + 2 * A_SLOTS + 12;
constexpr size_t F16_SMEM_BYTES = (size_t)F16_STAGES * F16_STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/;
static_assert(8 * F16_NUM_BARRIERS + 8 <= 512, "barrier block too small");
static_assert(F16_SMEM_BYTES <= 232448, "fp16 flavour exceeds the 227 KB of shared memory");
// ... and its rectangular form (method 5): 16 KiB of B + 16 KiB of raw A boxes + 6 KiB of raw B boxes, four stages.
constexpr int F16_RECT_STAGES = 4;
constexpr int F16_RECT_OFF_BRAW = F16_OFF_ARAW + RECT_RAW_A_BYTES;
constexpr int F16_RECT_STAGE_BYTES = OPERAND_BYTES + RECT_RAW_A_BYTES + RAW_BYTES;     // 38 KiB
static_assert(F16_RECT_STAGE_BYTES % 1024 == 0, "stage must keep the 1024-byte alignment of the swizzled tiles");
constexpr size_t F16_RECT_SMEM_BYTES = (size_t)F16_RECT_STAGES * F16_RECT_STAGE_BYTES + 1024 + 512;

struct HaloParams {
  const int4* rgs;        // [Dpad / 32] {stream | issue << 1, first channel, lag shift k0 - pre, byte offset in the operand's raw region}
  const int4* ktiles;     // [NK][2] {raw row of design row 0: stream 0, stream 1, begin | end << 8 | edge << 16, -t0}, {len0 - t0, len1 - t0, 0, 0}
  const int* out_idx;     // [Dpad] tile-space row -> row/col of G, -1 for padding
  // Rectangular mode (method 5): the A operand comes from its own row-group table (rect_ta pair tiles
  // of A rows against all nt pair tiles of B rows), every element (m, n) is written at
  // G[out_idx[n]][out_idx_a[m]] (no triangle), ntri = rect_ta * nt.
  int rect_ta;
  const int4* rgs_a;
  const int* out_idx_a;
  float* G;
  long long ldg;
  long long batch_stride; // 0: one Gram; else superchunk sc (= one segment) accumulates into G + sc * batch_stride
  int ntri;
  long long NK;
  int KS;
  int chunk_kt;
  long long total_steps;
  long long* prof;
  // fp16 x 3 flavour (F16 = true): the tensor maps cover the pre-split copy of the recording (one
  // 32-bit word per sample: fp16 hi | fp16 lo << 16 of the sample times its channel's power-of-two
  // scale); the epilogue multiplies element (m, n) by inv_scale_a[m] * inv_scale[n] (tile-space rows).
  const float* inv_scale;
  const float* inv_scale_a;
};

}  // namespace hk
template <>
__device__ __forceinline__ int rect_tiles_of<hk::HaloParams>(const hk::HaloParams& P) { return P.rect_ta; }
namespace hk {

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// L2 residency hints.  A superchunk of raw frames (~48 MB) is swept by every tile and should stay in
// L2; the partial tiles the epilogues red.add into G (ntri x 256 KB per superchunk, gigabytes) stream
// through it exactly once.  Measured at the cfg 4 width on one box: no hints 656 ms, raw boxes
// evict_last only 658, partials evict_first only 651, both 631 ms.  (The DRAM byte counts do not
// move -- 76 GB read + 28 GB written per 300 000 frames either way, ~6.5 % of the raw boxes miss L2
// for every superchunk size from 12 to 48 MB -- the gain is in L2 / crossbar time and power.)
#ifndef TD_HALO_L2_HINTS
#define TD_HALO_L2_HINTS 1
#endif
#ifndef TD_HALO_L2_HINT_TMA
#define TD_HALO_L2_HINT_TMA TD_HALO_L2_HINTS
#endif
#ifndef TD_HALO_L2_HINT_RED
#define TD_HALO_L2_HINT_RED TD_HALO_L2_HINTS
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void red_add_f32_hint(float* addr, float v, uint64_t policy) {
  asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(addr), "f"(v), "l"(policy) : "memory");
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// RECT = false: symmetric Gram (lower-triangle tile set, A and B from the same row groups).
// RECT = true:  rectangular product for method 5; tmap1 is then the A-side map (box 32 ch x 32 frames).
// F16 = false:  3xTF32 -- fp32 samples, hi = the raw bits (the tensor core ignores the low 13 mantissa
//               bits), lo = x - trunc(x), K = 8 per MMA.
// F16 = true:   3xFP16 -- the same 11 + 11 bit split (hi = fp16(x'), lo = fp16(x' - hi), x' = x times a
//               per-channel power of two that puts the channel's largest sample in [2^14, 2^15), so
//               that neither half leaves the fp16 range) done ONCE by split_rows_kernel; a converter
//               thread only packs two frames per register.  K = 16 per MMA: half the tensor-pipe
//               cycles of 3xTF32 for the same products hi*hi + hi*lo + lo*hi in fp32 accumulators.
template <bool RECT, bool F16>
__global__ void __launch_bounds__(F16 ? NUM_THREADS_F16 : NUM_THREADS, 1)
lagged_gram_halo_kernel(const __grid_constant__ CUtensorMap tmap0,
                        const __grid_constant__ CUtensorMap tmap1, const HaloParams P) {
  constexpr uint32_t IDESC = F16 ? ptx::umma_idesc_f16(256, 128) : ptx::umma_idesc_tf32(256, 128);
  constexpr int STAGES = F16 ? (RECT ? F16_RECT_STAGES : F16_STAGES) : hk::STAGES;
  constexpr int NUM_BARRIERS = 2 * STAGES + 2 * A_SLOTS + 12;
  constexpr int STAGE_BYTES = F16 ? (RECT ? F16_RECT_STAGE_BYTES : F16_STAGE_BYTES) : (RECT ? RECT_STAGE_BYTES : hk::STAGE_BYTES);
  constexpr int OFF_ARAW = F16 ? F16_OFF_ARAW : hk::OFF_ARAW;
  constexpr int OFF_BRAW = F16 ? (RECT ? F16_RECT_OFF_BRAW : F16_OFF_BRAW) : (RECT ? RECT_OFF_BRAW : hk::OFF_BRAW);
  constexpr int OFF_BLO = F16 ? F16_OFF_BLO : hk::OFF_BLO;
  constexpr int EPI_WARP0 = F16 ? 12 : 8;
  // Register budget per warpgroup (regs per thread; the launch allocates 128 x 512 resp. 96 x 640):
  //   3xTF32:  72 + 104 + 168 + 168 = 512        3xFP16:  40 + 72 + 72 + 144 + 144 = 472 <= 480
  constexpr int REGS_WG0 = F16 ? 48 : 72, REGS_CONV = F16 ? 72 : 104, REGS_EPI = F16 ? 144 : 168;
  using ts::HalfRing;
  using ts::lo_bits;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // raw boxes landed (local)
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };           // MMAs of the stage retired (both issuers)
  auto ready_bar = [&](int j) { return bar_base + 8u * (2 * STAGES + j); };       // leader: A slot + B hi/lo of 16 frames built in both CTAs
  auto aempty_bar = [&](int j) { return bar_base + 8u * (2 * STAGES + A_SLOTS + j); };   // MMAs reading the slot retired
  auto tfull_bar = [&](int h, int r) { return bar_base + 8u * (2 * STAGES + 2 * A_SLOTS + 3 * h + r); };
  auto tempty_bar = [&](int h, int r) { return bar_base + 8u * (2 * STAGES + 2 * A_SLOTS + 6 + 3 * h + r); };
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
      ptx::mbar_init(empty_bar(s), 2);
    }
    for (int j = 0; j < A_SLOTS; ++j) {
      ptx::mbar_init(ready_bar(j), 2 * NUM_CONV);
      ptx::mbar_init(aempty_bar(j), 2);
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

  const int wg = warp >> 2;
  if (wg == 0) {
    ptx::setmaxnreg_dec<REGS_WG0>();
    if (warp == 0) {
      // ===================== TMA producer: raw boxes of this CTA's A and B row groups =============
      WorkIter wi(P, worker, n_workers);
      WorkItem it;
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t keep_policy = l2_policy_evict_last();
      while (wi.next(P, it)) {
        const int ta = (it.tm * 2 + (int)rank) * RG_PER_CTA;
        const int tb = (it.tn * 2 + (int)rank) * RG_PER_CTA;
        // per row group, packed into one register: channel << 17 | (lag shift & 0x7fff) << 2 | issue << 1 | stream
        uint32_t rg[2 * RG_PER_CTA];
        const int4* rgs_of_a = RECT ? P.rgs_a : P.rgs;
        uint32_t tx_bytes = 0;
#pragma unroll
        for (int i = 0; i < 2 * RG_PER_CTA; ++i) {
          const int4 g = __ldg(i < RG_PER_CTA ? &rgs_of_a[ta + i] : &P.rgs[tb + (i & (RG_PER_CTA - 1))]);
          rg[i] = ((uint32_t)g.y << 17) | (((uint32_t)g.z & 0x7fffu) << 2) | ((uint32_t)g.x & 3u);
          if (g.x & 2) tx_bytes += (RECT && i < RG_PER_CTA) ? ABOX_BYTES : BOX_BYTES;
        }
        int4 kn = __ldg(&P.ktiles[2 * it.kt_begin]);
        for (long long kt = it.kt_begin; kt < it.kt_end; ++kt) {
          const int4 kc = kn;
          if (kt + 1 < it.kt_end) kn = __ldg(&P.ktiles[2 * (kt + 1)]);
          if constexpr (F16) ptx::mbar_wait_relaxed<TD_F16_SLEEP_NS>(empty_bar(stage), phase ^ 1u);   // the ring is normally full: this warp only waits
          else ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t fb = full_bar(stage);
            ptx::mbar_expect_tx(fb, tx_bytes);
            const uint32_t sbase = smem_base + stage * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 2 * RG_PER_CTA; ++i) {
              if (rg[i] & 2u) {
                const int q = (int)(rg[i] & 1u);
                const int c0 = (int)(rg[i] >> 17);
                const int d0 = ((int)(rg[i] << 15)) >> 17;          // sign-extended 15-bit lag shift
                if (RECT && i < RG_PER_CTA) {
                  if (TD_HALO_L2_HINT_TMA) tma_load_2d_hint(sbase + OFF_ARAW + i * ABOX_BYTES, &tmap1, fb, c0, kc.x + d0, keep_policy);
                  else tma_load_2d(sbase + OFF_ARAW + i * ABOX_BYTES, &tmap1, fb, c0, kc.x + d0);
                } else {
                  const uint32_t dst = sbase + (i < RG_PER_CTA ? OFF_ARAW : OFF_BRAW) + (i & (RG_PER_CTA - 1)) * BOX_BYTES;
                  if (TD_HALO_L2_HINT_TMA) tma_load_2d_hint(dst, (!RECT && q) ? &tmap1 : &tmap0, fb, c0, (q ? kc.y : kc.x) + d0, keep_policy);
                  else tma_load_2d(dst, (!RECT && q) ? &tmap1 : &tmap0, fb, c0, (q ? kc.y : kc.x) + d0);
                }
              }
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp <= 2 && rank == 0) {
      // ===================== MMA issuers (pair leader): warp 1 -> half 0, warp 2 -> half 1 =========
      const int h = warp - 1;
      WorkIter wi(P, worker, n_workers);
      WorkItem it;
      int stage = 0;
      uint32_t kcount = 0;                       // k-tiles issued so far: A slots 2*kcount, 2*kcount + 1
      HalfRing ring;
      while (wi.next(P, it)) {
        int next_range = __ldg(&P.ktiles[2 * it.kt_begin]).z;
        for (long long c0 = it.kt_begin; c0 < it.kt_end; c0 += P.chunk_kt) {
          const long long c1 = min(it.kt_end, c0 + (long long)P.chunk_kt);
          ring.next_chunk(h);
          for (long long kt = c0; kt < c1; ++kt, ++kcount) {
            const int range = next_range;
            if (kt + 1 < it.kt_end) next_range = __ldg(&P.ktiles[2 * (kt + 1)]).z;
            const int ks0 = (range & 0xff) >> 3, ks1 = (((range >> 8) & 0xff) + 7) >> 3;
            const bool whole = (ks0 == 0) && (ks1 == BK / 8);
            const uint32_t sbase = smem_base + stage * STAGE_BYTES;
            const uint32_t first_acc = (kt > c0) ? 1u : 0u;
            const int r = ring.r;
            const uint32_t d = tmem_base + (uint32_t)r * ACC_COLS;
            if (kt == c0) ptx::mbar_wait(tempty_bar(1 - h, r), ring.drained_parity(h));
            const uint64_t b_hi = ptx::umma_desc_k_sw128(sbase + OFF_BHI + h * (OPERAND_BYTES / 2));
            // (fp16 x 3: the lo halves are bytes [64, 128) of the same rows -- 4 descriptor units further)
            const uint64_t b_lo = F16 ? b_hi + (uint64_t)(F16_OFF_BLO >> 4)
                                      : ptx::umma_desc_k_sw128(sbase + OFF_BLO + h * (OPERAND_BYTES / 2));
            if constexpr (F16) {
              // One A slot = one k-tile (32 frames = two K = 16 steps: hi columns [8s, 8s + 8), lo columns
              // [16 + 8s, ...); 32 bytes of a B row per step).  Frames outside the call's row range are
              // zero in A, so partial k-tiles need no skipping.
              const uint32_t slot = kcount & (A_SLOTS - 1);
              if constexpr (TD_F16_ISSUER_SLEEP_NS > 0) ptx::mbar_wait_relaxed<TD_F16_ISSUER_SLEEP_NS>(ready_bar(slot), (kcount >> 2) & 1u);
              else ptx::mbar_wait_lean(ready_bar(slot), (kcount >> 2) & 1u);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint32_t a_base = tmem_base + A_COL0 + slot * A_SLOT_COLS;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t adv = (uint64_t)(ks * 2);
                  ptx::mma_f16_ts_pair(d, a_base + 8 * ks, b_hi + adv, IDESC, ks ? 1u : first_acc);
                  ptx::mma_f16_ts_pair(d, a_base + 8 * ks, b_lo + adv, IDESC, 1u);
                  ptx::mma_f16_ts_pair(d, a_base + 16 + 8 * ks, b_hi + adv, IDESC, 1u);
                }
                ptx::mma_commit<2>(aempty_bar(slot));
                ptx::mma_commit<2>(empty_bar(stage));
                if (kt == c1 - 1) ptx::mma_commit<2>(tfull_bar(h, r));
              }
              __syncwarp();
            } else {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t use = 2u * kcount + hh;
              const uint32_t slot = use & (A_SLOTS - 1);
              ptx::mbar_wait(ready_bar(slot), (use >> 2) & 1u);
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
                ptx::mma_commit<2>(aempty_bar(slot));                        // A slot reusable
                if (hh == 1) {
                  ptx::mma_commit<2>(empty_bar(stage));                      // stage reusable in both CTAs
                  if (kt == c1 - 1) ptx::mma_commit<2>(tfull_bar(h, r));
                }
              }
              __syncwarp();
            }
            }
            if (++stage == STAGES) stage = 0;
          }
        }
      }
    }
  } else if (!F16 && wg == 1) {
    // ===================== converters: raw boxes -> A in TMEM (hi | lo), B hi / lo tiles ==========
    ptx::setmaxnreg_dec<REGS_CONV>();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                    // A row (= TMEM lane) and B row of this thread
    const int kk = lane >> 3, cc = lane & 7;             // lag and channel of the row inside its row group
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t ready_leader0 = ptx::mapa(ready_bar(0), 0);
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + A_COL0;
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t kcount = 0;
    while (wi.next(P, it)) {
      const int4 ga = __ldg(&(RECT ? P.rgs_a : P.rgs)[(it.tm * 2 + (int)rank) * RG_PER_CTA + quad]);
      const int4 gb = __ldg(&P.rgs[(it.tn * 2 + (int)rank) * RG_PER_CTA + quad]);
      // RECT: lane = channel of a 32-channel, one-lag row group (32 consecutive words per frame)
      const uint32_t a_rd = OFF_ARAW + (uint32_t)ga.w + (RECT ? (uint32_t)lane * 4u : (uint32_t)(kk * BOX_CH + cc) * 4u);
      const uint32_t b_rd = OFF_BRAW + (uint32_t)gb.w + (uint32_t)(kk * BOX_CH + cc) * 4u;
      constexpr uint32_t A_FRAME_BYTES = RECT ? ABOX_CH * 4 : BOX_CH * 4;
      const int a_d = ga.z + (RECT ? 0 : kk), b_d = gb.z + kk;        // lag shift k - pre of this thread's rows
      const int a_q = ga.x & 1, b_q = gb.x & 1;
      int4 kn0 = __ldg(&P.ktiles[2 * it.kt_begin]);
      int4 kn1 = __ldg(&P.ktiles[2 * it.kt_begin + 1]);
      for (long long kt = it.kt_begin; kt < it.kt_end; ++kt, ++kcount) {
        const int4 k0 = kn0, k1 = kn1;
        if (kt + 1 < it.kt_end) {
          kn0 = __ldg(&P.ktiles[2 * (kt + 1)]);
          kn1 = __ldg(&P.ktiles[2 * (kt + 1) + 1]);
        }
        const int begin = k0.z & 0xff, end = (k0.z >> 8) & 0xff;
        const bool edge = (k0.z >> 16) != 0;
        const bool partial = begin > 0 || end < BK;
        // element e of the k-tile (design row t0 + e) at lag shift d reads frame t0 + e + d of the
        // segment: inside the trial iff -t0 <= e + d < len - t0
        const int a_lo_e = k0.w - a_d, a_hi_e = (a_q ? k1.y : k1.x) - a_d;
        const int b_lo_e = k0.w - b_d, b_hi_e = (b_q ? k1.y : k1.x) - b_d;
        const uint32_t sbase = smem_base + stage * STAGE_BYTES;
        const uint32_t bhi = sbase + OFF_BHI + row_off, blo = sbase + OFF_BLO + row_off;
        ptx::mbar_wait(full_bar(stage), phase);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t use = 2u * kcount + hh;
          const uint32_t slot = use & (A_SLOTS - 1);
          uint32_t hi[16], lo[16], bv[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            hi[e] = lds32(sbase + a_rd + (uint32_t)(hh * 16 + e) * A_FRAME_BYTES);
            bv[e] = lds32(sbase + b_rd + (uint32_t)((hh * 16 + e) * BOX_CH * 4));
          }
          if (edge) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int t = hh * 16 + e;
              if (t < a_lo_e || t >= a_hi_e) hi[e] = 0u;
              if (t < b_lo_e || t >= b_hi_e) bv[e] = 0u;
            }
          }
          if (partial) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int t = hh * 16 + e;
              if (t < begin || t >= end) hi[e] = 0u;
            }
          }
          // B: hi tile (the raw bits) and lo tile, K-major with the 128-byte swizzle: frame t of a row
          // sits in 16-byte chunk (t/4) ^ (row%8); lane-per-row 16-byte stores are conflict free.
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t off = ((uint32_t)(hh * 4 + i) ^ sw) << 4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bhi + off),
                         "r"(bv[4 * i]), "r"(bv[4 * i + 1]), "r"(bv[4 * i + 2]), "r"(bv[4 * i + 3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blo + off),
                         "r"(lo_bits(bv[4 * i])), "r"(lo_bits(bv[4 * i + 1])), "r"(lo_bits(bv[4 * i + 2])),
                         "r"(lo_bits(bv[4 * i + 3]))
                         : "memory");
          }
          ptx::fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor-core proxy
#pragma unroll
          for (int e = 0; e < 16; ++e) lo[e] = lo_bits(hi[e]);
          ptx::mbar_wait(aempty_bar(slot), ((use >> 2) & 1u) ^ 1u);
          ptx::tc_fence_after();
          ptx::tmem_st16(tlane + slot * A_SLOT_COLS, hi);
          ptx::tmem_st16(tlane + slot * A_SLOT_COLS + 16, lo);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(ready_leader0 + 8u * slot);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (F16 && wg <= 2) {
    // ===================== fp16 x 3 converters: two warpgroups, k-tiles alternate between them ==========
    // Every raw word is (hi | lo << 16) of one sample (split_rows_kernel); converting means packing
    // two frames per 32-bit register.  A thread owns row `row` of the A tile (= its TMEM lane) and row
    // `row` of the B tile, like the 3xTF32 converter, but a k-tile is only 768 tensor-pipe clocks now
    // and one warp's chain (waits, 64 loads, 64 packs, stores, fences, arrive: ~1000 clocks, latency
    // bound) does not fit -- so warpgroup 1 converts the even k-tiles and warpgroup 2 the odd ones.
    //   A: hi halves of frames [16h, 16h + 16) = TMEM columns [8h, 8h + 8) of the slot, lo halves =
    //      columns [16 + 8h, ...);
    //   B: frame t of a row sits in 16-byte chunk (t/8) ^ (row%8) of the hi tile and of the lo tile.
    ptx::setmaxnreg_dec<REGS_CONV>();
    static_assert(A_SLOTS == 4, "slot / parity below are bit fields of the k-tile count");
    const int my_par = wg - 1;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int kk = lane >> 3, cc = lane & 7;
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t ready_leader0 = ptx::mapa(ready_bar(0), 0);
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + A_COL0;
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    uint32_t kcount = 0;                                 // k-tiles of all earlier work items
    while (wi.next(P, it)) {
      const int4 ga = __ldg(&(RECT ? P.rgs_a : P.rgs)[(it.tm * 2 + (int)rank) * RG_PER_CTA + quad]);
      const int4 gb = __ldg(&P.rgs[(it.tn * 2 + (int)rank) * RG_PER_CTA + quad]);
      const uint32_t a_rd = OFF_ARAW + (uint32_t)ga.w + (RECT ? (uint32_t)lane * 4u : (uint32_t)(kk * BOX_CH + cc) * 4u);
      const uint32_t b_rd = OFF_BRAW + (uint32_t)gb.w + (uint32_t)(kk * BOX_CH + cc) * 4u;
      constexpr uint32_t A_FRAME_BYTES = RECT ? ABOX_CH * 4 : BOX_CH * 4;
      const int a_d = ga.z + (RECT ? 0 : kk), b_d = gb.z + kk;        // lag shift k - pre of this thread's rows
      const int a_q = ga.x & 1, b_q = gb.x & 1;
      const uint32_t first = (uint32_t)((my_par - (int)kcount) & 1);  // offset of this group's first k-tile
      long long kt = it.kt_begin + first;
      uint32_t g = kcount + first;                                    // running k-tile count of kt
      kcount += (uint32_t)(it.kt_end - it.kt_begin);
      int kz_next = (kt < it.kt_end) ? __ldg(&P.ktiles[2 * kt].z) : 0;
      uint32_t stage = g % (uint32_t)STAGES, phase = (g / (uint32_t)STAGES) & 1u;   // raw ring position of kt
      for (; kt < it.kt_end; kt += 2, g += 2) {
        const int kz = kz_next;                           // begin | end << 8 | edge << 16 of this k-tile
        if (kt + 2 < it.kt_end) kz_next = __ldg(&P.ktiles[2 * (kt + 2)].z);
        const uint32_t slot = g & 3u, par = (g >> 2) & 1u;             // A slot and its parity
        const uint32_t sbase = smem_base + stage * STAGE_BYTES;
        const bool special = kz != (BK << 8);             // anything but a whole k-tile in the middle of a trial
        int begin = 0, end = BK, a_lo_e = 0, a_hi_e = BK, b_lo_e = 0, b_hi_e = BK;
        bool edge = false, partial = false;
        if (special) {
          begin = kz & 0xff;
          end = (kz >> 8) & 0xff;
          edge = (kz >> 16) != 0;
          partial = begin > 0 || end < BK;                // the row range is cut on the A side only
          const int neg_t0 = __ldg(&P.ktiles[2 * kt].w);  // inside the trial iff lo_e <= e < hi_e
          const int2 len_t0 = *reinterpret_cast<const int2*>(&P.ktiles[2 * kt + 1]);
          a_lo_e = neg_t0 - a_d;
          b_lo_e = neg_t0 - b_d;
          a_hi_e = (a_q ? len_t0.y : len_t0.x) - a_d;
          b_hi_e = (b_q ? len_t0.y : len_t0.x) - b_d;
        }
        const uint32_t bhi = sbase + OFF_BHI + row_off;
        ptx::mbar_wait_lean(full_bar(stage), phase);
        ptx::mbar_wait_lean(aempty_bar(slot), par ^ 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                  // 16 frames at a time keeps the register footprint low
          uint32_t wa[16], wb[16], p2[8], q2[8];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            wa[e] = lds32(sbase + a_rd + (uint32_t)(hh * 16 + e) * A_FRAME_BYTES);
            wb[e] = lds32(sbase + b_rd + (uint32_t)((hh * 16 + e) * BOX_CH * 4));
          }
          if (special) {
            if (edge) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int t = hh * 16 + e;
                if (t < a_lo_e || t >= a_hi_e) wa[e] = 0u;
                if (t < b_lo_e || t >= b_hi_e) wb[e] = 0u;
              }
            }
            if (partial) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int t = hh * 16 + e;
                if (t < begin || t >= end) wa[e] = 0u;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            p2[j] = __byte_perm(wa[2 * j], wa[2 * j + 1], 0x5410);       // hi halves of frames 2j, 2j + 1
            q2[j] = __byte_perm(wa[2 * j], wa[2 * j + 1], 0x7632);       // lo halves
          }
          ptx::tmem_st8(tlane + slot * A_SLOT_COLS + 8 * hh, p2);
          ptx::tmem_st8(tlane + slot * A_SLOT_COLS + 16 + 8 * hh, q2);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            p2[j] = __byte_perm(wb[2 * j], wb[2 * j + 1], 0x5410);
            q2[j] = __byte_perm(wb[2 * j], wb[2 * j + 1], 0x7632);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t off = ((uint32_t)(hh * 2 + i) ^ sw) << 4;
            const uint32_t off_lo = ((uint32_t)(4 + hh * 2 + i) ^ sw) << 4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bhi + off), "r"(p2[4 * i]),
                         "r"(p2[4 * i + 1]), "r"(p2[4 * i + 2]), "r"(p2[4 * i + 3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bhi + off_lo), "r"(q2[4 * i]),
                         "r"(q2[4 * i + 1]), "r"(q2[4 * i + 2]), "r"(q2[4 * i + 3])
                         : "memory");
          }
        }
        ptx::fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor-core proxy
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(ready_leader0 + 8u * slot);
        stage += 2;
        if (stage >= (uint32_t)STAGES) { stage -= STAGES; phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: TMEM region -> fp32 registers -> red.add =====================
    ptx::setmaxnreg_inc<REGS_EPI>();
    const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - EPI_WARP0) >> 2;  // which 128-column half of the pair tile
    const uint32_t tempty_leader0 = ptx::mapa(tempty_bar(half, 0), 0);
    WorkIter wi(P, worker, n_workers);
    WorkItem it;
    HalfRing ring;
    float acc[NACC];
    const uint64_t stream_policy = l2_policy_evict_first();
    while (wi.next(P, it)) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
      for (long long c0 = it.kt_begin; c0 < it.kt_end; c0 += P.chunk_kt) {
        ring.next_chunk(half);
        const int r = ring.r;
        if constexpr (F16) ptx::mbar_wait_relaxed<TD_F16_SLEEP_NS>(tfull_bar(half, r), ring.n);     // idle for most of a chunk
        else ptx::mbar_wait(tfull_bar(half, r), ring.n);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)r * ACC_COLS;
        if constexpr (F16) {                 // 144 registers: 16 columns at a time
#pragma unroll
          for (int j = 0; j < NACC / 16; ++j) {
            float v[16];
            ptx::tmem_ld16(taddr + j * 16, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[j * 16 + i] += v[i];
          }
        } else {
#pragma unroll
          for (int j = 0; j < NACC / 32; ++j) {
            float v[32];
            ptx::tmem_ld32(taddr + j * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(tempty_leader0 + 8u * r);
      }
      // Tile (tm, tn), tm >= tn: element (m, n), n <= m in TILE space, is added at G[n_out][m_out];
      // that is either triangle of G (the tile-space row order is not the design's column order),
      // fold_symmetric_kernel adds the two triangles afterwards.  The 32 lanes of a warp are the 32
      // rows of one row group (4 lags x 8 channels): 4 runs of 8 consecutive addresses.  Accumulator
      // column i of half h is B row h*64 + i%64 of CTA i/64.
      const int m_row = (it.tm * 2 + (int)rank) * TILE_ROWS + quad * 32 + lane;
      const int m_out = __ldg(&(RECT ? P.out_idx_a : P.out_idx)[m_row]);
      const bool all_pairs = RECT || it.tm > it.tn;
      float* Gout = P.G + (it.kt_begin / P.KS) * P.batch_stride;      // batched: one Gram per superchunk
      float inv_m = 1.0f;
      if constexpr (F16) inv_m = __ldg(&(RECT ? P.inv_scale_a : P.inv_scale)[m_row]);
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        const int n_row = it.tn * 256 + (i >> 6) * TILE_ROWS + half * 64 + (i & 63);
        const int n_out = __ldg(&P.out_idx[n_row]);
        if (m_out >= 0 && n_out >= 0 && (all_pairs || n_row <= m_row)) {
          float v = acc[i];
          if constexpr (F16) v = (v * inv_m) * __ldg(&P.inv_scale[n_row]);     // exact: powers of two
          if (TD_HALO_L2_HINT_RED) red_add_f32_hint(Gout + (long long)n_out * P.ldg + m_out, v, stream_policy);
          else atomicAdd(Gout + (long long)n_out * P.ldg + m_out, v);
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc<2>(tmem_base, TMEM_COLS);
}

}  // namespace hk
