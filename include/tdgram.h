/*
 * tdgram.h -- C ABI of libtdgram.so, the B200 (sm_100a) implementation of the dense
 * decoding hot path of google/telluride_decoding.
 *
 * The reference has no FFI seam: its hot path is numpy called from Python
 * (telluride_decoding/brain_model.py:429-481, cca.py:311-369, infer_decoder.py:288-328,
 * result_store.py:253-338).  Each entry point below names the reference code it replaces.
 * The Python package telluride_decoding_b200 binds these with ctypes
 * (telluride_decoding_b200/_lib.py); INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - Every function returns 0 on success or a negative td_status; td_last_error() gives a
 *     thread-local message.  Nothing throws, nothing allocates caller-visible memory.
 *   - Pointers named *_dev are CUDA device pointers (they come from torch tensors that crossed
 *     by DLPack); `segments` tables and lambda lists are small HOST arrays.
 *   - All work is enqueued on `cuda_stream` (a cudaStream_t cast to void*, NULL = default
 *     stream).  Functions do not synchronise unless documented.
 *   - Matrices are row-major float32 unless stated.
 *
 * Lag (temporal context) semantics -- brain_data.py:425-483.  A *stream* is a raw
 * [rows, channels] array cut into *segments* (one per trial file and per
 * `initial_batch_size` chunk).  Design row t of segment s is
 *     X_s[t, k*C + c] = x_s[t - pre + k, c]   for k in [0, pre+1+post),  0 outside [0, len_s)
 * and only the first keep[s] rows of each segment are used (`drop_remainder`,
 * brain_data.py:369-370).  The lagged design is never written to memory.
 */
#ifndef TDGRAM_H_
#define TDGRAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TD_VERSION 100

typedef enum td_status {
  TD_OK = 0,
  TD_ERR_INVALID = -1,      /* bad argument (message says which) */
  TD_ERR_WORKSPACE = -2,    /* workspace too small */
  TD_ERR_CUDA = -3,         /* CUDA runtime / driver error */
  TD_ERR_SOLVER = -4,       /* cuSOLVER / cuBLAS failure or singular system */
  TD_ERR_UNSUPPORTED = -5   /* shape outside what the kernels handle */
} td_status;

/* One raw input stream with its lag spec.  seg_start/seg_len are HOST arrays of n_seg entries. */
typedef struct td_stream {
  const float* data_dev;     /* [rows, channels], row stride ld elements */
  int64_t ld;
  int32_t channels;
  int32_t pre;               /* pre_context  (brain_data.py:104) */
  int32_t post;              /* post_context (brain_data.py:105) */
  const int64_t* seg_start;  /* first raw row of segment i in data_dev */
  const int64_t* seg_len;    /* raw rows in segment i: zero padding begins at seg_len[i] */
} td_stream;

const char* td_last_error(void);
int td_version(void);
/* Number of SMs / compute capability of the current device (for the bench's roofline). */
int td_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Releases what the library keeps between calls (the calling thread's timing events, cuSOLVER /
 * cuBLAS handles, their side streams and events).  The caller's streams must be idle; the library
 * stays usable (everything is recreated on demand). */
int td_shutdown(void);

/* ---------------------------------------------------------------------------------------
 * Lagged Gram.  Replaces the minibatch loop `sum_xtx += np.matmul(x.T, x)`
 * (brain_model.py:437) and the three products cov_xx/cov_yy/cov_xy of cca.py:325-327.
 *
 * The design is Z = [X_0 | X_1 | ...] over n_streams (1 or 2) lagged streams sharing the
 * segment table; G = Z'Z, [Dz, Dz] with Dz = sum_q channels_q * (pre_q+1+post_q), full
 * symmetric, row stride ldg.  method 0: tcgen05 implicit-Hankel kernel on split operands (every sample
 * x = hi + lo with 11 + 11 significant bits, products hi*hi + hi*lo + lo*hi, TMEM fp32 accumulation
 * flushed every `chunk_rows` frames into fp32 registers: fp32-equivalent, relative Frobenius error
 * ~1e-6); it picks between the flavours that can also be forced:
 *   6  CTA pairs, 3xFP16 (kind::f16, K = 16 per MMA): the method-4 kernel on a PRE-SPLIT copy of the
 *      recording -- one pass finds every channel's largest sample, a second writes one 32-bit word per
 *      sample, fp16(x') | fp16(x' - fp16(x')) << 16 with x' = x * 2^e_c (the power of two that puts the
 *      channel's peak in [2^14, 2^15), undone exactly in the epilogue), so neither half leaves the
 *      fp16 range whatever the unit of the recording.  Same split as 3xTF32 with round-to-nearest hi
 *      halves (measured 7e-7 against 2.3e-6), half the tensor-pipe cycles; on a power-capped B200
 *      1.45 - 1.5x the frames per second of method 4.  Default for designs of >= 512 columns.
 *      Workspace: tables + the pre-split copy (4 bytes per touched sample).
 *   4  CTA pairs, 3xTF32 (kind::tf32, K = 8); the lagged operand is built ON CHIP: one raw [32 + halo frames, 8 channels] TMA box
 *      per 4-lag x 8-channel row group straight from the caller's time-major array, converter warps
 *      read every row at its own lag offset, rebuild the zero padding of the trial edges with masks and
 *      write A into tensor memory and B into swizzled shared memory.  Workspace: tables only (plus
 *      one padded copy of the touched rows when the row stride or base address is not a multiple of
 *      16 bytes).
 *   3  CTA pairs, A operand in tensor memory, operands fetched as (lag, channel-block) TMA boxes from
 *      four shifted channel-major copies of the recording staged in the workspace (round-1 kernel).
 *   2  both operands from shared memory (the only flavour for cta_group 1, i.e. narrow designs).
 *   5  (opt-in, one stream) the block-Toeplitz structure of a lagged Gram: only the first block row
 *      C(d)[c1,c2] = sum_t x[t-pre,c1] x[t-pre+d,c2] is contracted on the tensor cores (the method-4
 *      kernel in rectangular mode, L/2 times fewer MMAs), every other block follows along the lag
 *      diagonals from C(d) plus edge sums over the first / last k1 frames of each segment.  Same
 *      split arithmetic (the fp16 x 3 flavour; reserved[1] = 1: 3xTF32) and row semantics; the dense
 *      flavours stay the default and the reference point of the roofline.
 * method 1: fp32 CUDA-core kernel kept as an on-device cross-check.
 * Workspace: td_lagged_gram_workspace_bytes bytes of device memory.
 * ------------------------------------------------------------------------------------- */
typedef struct td_gram_options {
  int32_t method;        /* 0 = tcgen05 split-operand kernel (default), 1 = fp32 SIMT cross-check, 2 / 3 / 4 / 6 = force a dense tcgen05 flavour, 5 = block-Toeplitz */
  int32_t cta_group;     /* 0 = auto, 1 = 128x128 tiles per CTA, 2 = CTA-pair 256x256 tiles */
  int32_t chunk_rows;    /* TMEM flush period in frames (0 = default: 256 for method 6, else 128) */
  int32_t superchunk_rows; /* L2 working-set period in frames (0 = auto) */
  int32_t max_ctas;      /* 0 = all SMs */
  int32_t reserved[3];   /* reserved[0] = 1: record CUDA events around the Gram kernel; reserved[1] = 1: method 5 contracts in 3xTF32 */
} td_gram_options;

int td_lagged_gram_workspace_bytes(const td_stream* streams, int n_streams,
                                   const int64_t* seg_first, const int64_t* seg_keep, int n_seg,
                                   const td_gram_options* opt, size_t* bytes);
/* The same for td_lagged_gram_batched_f32 (method 5 keeps one block row per segment in the workspace). */
int td_lagged_gram_batched_workspace_bytes(const td_stream* streams, int n_streams,
                                           const int64_t* seg_first, const int64_t* seg_keep, int n_seg,
                                           const td_gram_options* opt, size_t* bytes);
/* Design rows [seg_first[s], seg_keep[s]) of every segment are accumulated (seg_first NULL = 0):
 * a prefix for drop_remainder, a sub-range when a trial's frames are shared between GPUs (the
 * lag halo is read from the neighbouring frames of the same segment, not zero padded). */
int td_lagged_gram_f32(const td_stream* streams, int n_streams,
                       const int64_t* seg_first, const int64_t* seg_keep, int n_seg,
                       float* G_dev, int64_t ldg, const td_gram_options* opt,
                       void* workspace_dev, size_t workspace_bytes, void* cuda_stream);
/* One Gram PER SEGMENT from one launch (methods 6 (default), 4 or 5): G_dev [n_seg][Dz][ldg].  What leave-one-trial-out
 * needs (regression.py:197-214, infer.py: each held-out trial's fold = total - trial): every
 * segment must span the same, non-zero number of 32-frame k-tiles. */
int td_lagged_gram_batched_f32(const td_stream* streams, int n_streams,
                               const int64_t* seg_first, const int64_t* seg_keep, int n_seg,
                               float* G_dev, int64_t ldg, const td_gram_options* opt,
                               void* workspace_dev, size_t workspace_bytes, void* cuda_stream);
/* out_dev[i] (double) = sum over b < batches of in_dev[b * stride + i]: the total Gram from the
 * per-segment Grams. */
int td_sum_batches_f32(const float* in_dev, int64_t n, int64_t stride, int batches, double* out_dev,
                       void* cuda_stream);
/* Device time (ms) of the most recent tcgen05 Gram kernel launched by this thread with
 * opt->reserved[0] = 1 (CUDA events on the caller's stream; synchronises on the stop event).
 * Negative if none was recorded.  For the bench's roofline line. */
float td_last_gram_kernel_ms(void);

/* ---------------------------------------------------------------------------------------
 * Thin products against the un-lagged targets and the ones column.  Replaces
 * `sum_xty += x.T @ y`, `sum_x += sum(x)` and the hstack'ed ones column
 * (brain_model.py:436-439), `sum_x`/`sum_y` of cca.py:328-329.
 *   V = [v_0 .. v_{Q-2}, 1]  (v = thin stream without lags, may be NULL: then V = [1])
 *   xtv_dev [Dx, Q] double  = X'V        vtv_dev [Q, Q] double = V'V   (V'V[Q-1,Q-1] = row count)
 * over design rows [seg_first[s], seg_keep[s]) (seg_first may be NULL = 0).
 * Both are zeroed by the call.  Q = (v ? v->channels : 0) + 1 <= 16.
 * ------------------------------------------------------------------------------------- */
int td_lagged_thin_f64(const td_stream* x, const td_stream* v,
                       const int64_t* seg_first, const int64_t* seg_keep, int n_seg,
                       double* xtv_dev, double* vtv_dev, void* cuda_stream);

/* The same per segment: xtv_dev [n_seg][Dx][Q], vtv_dev [n_seg][Q][Q]. */
int td_lagged_thin_per_segment_f64(const td_stream* x, const td_stream* v,
                                   const int64_t* seg_first, const int64_t* seg_keep, int n_seg,
                                   double* xtv_dev, double* vtv_dev, void* cuda_stream);

/* Ledoit-Wolf shrinkage statistics (brain_model.py:438-443, :457-465), from the implicit design.
 * Segments here are the PIECES of the minibatches (a minibatch of the interleaved stream is a
 * contiguous row range of each file); seg_slot[i] (HOST) is the minibatch piece i belongs to.
 *   td_lagged_column_sums_f64  sums_dev [n_slots][Dx], rows_dev [n_slots] double: np.sum(x, axis=0) and
 *                              the row count of every minibatch
 *   td_lagged_lw_stat_f64      out_dev[0] = sum_t (sum_i (X[t,i] - mean_dev[slot(t)][i])^2)^2, i.e. the sum
 *                              of all entries of the reference's sum_x2tx2 (the only thing it uses),
 *                              mean_dev [n_slots][Dx] float32 = the running mean after each minibatch */
int td_lagged_column_sums_f64(const td_stream* x, const int64_t* seg_first, const int64_t* seg_keep,
                              const int32_t* seg_slot, int n_seg, int n_slots, double* sums_dev,
                              double* rows_dev, void* cuda_stream);
int td_lagged_lw_stat_f64(const td_stream* x, const int64_t* seg_first, const int64_t* seg_keep,
                          const int32_t* seg_slot, int n_seg, int n_slots, const float* mean_dev,
                          double* out_dev, void* cuda_stream);

/* Gram over the row ranges [seg_first[s], seg_keep[s]) of each segment (seg_first NULL = 0), on
 * CUDA cores in fp32 with float64 carry.  Used for the small leave-one-out corrections: the
 * < batch_size rows that `drop_remainder` (brain_data.py:369-370) removes from a fold's training
 * stream are a suffix of some segments, and their Gram is subtracted from the all-rows Gram. */
int td_lagged_gram_range_f32(const td_stream* streams, int n_streams, const int64_t* seg_first,
                             const int64_t* seg_keep, int n_seg, float* G_dev, int64_t ldg,
                             void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * Forward pass.  Replaces Keras Dense on the lagged batch (brain_model.py:340-341) and one half
 * of BrainCcaLayer.call (cca.py:157-158):
 *   out[out_row[s] + t, :] = (X_s[t, :] - mean) W + bias      t < keep[s]
 * W [Dx, Dy] row-major, bias [Dy] or NULL, mean [Dx] or NULL, out row stride ldo, Dy <= 64.
 * out_row is a HOST array.
 * ------------------------------------------------------------------------------------- */
int td_lagged_predict_f32(const td_stream* x, const int64_t* seg_keep, const int64_t* out_row,
                          int n_seg, const float* W_dev, const float* bias_dev,
                          const float* mean_dev, int Dy, float* out_dev, int64_t ldo,
                          void* cuda_stream);

/* The same with ONE weight matrix per segment: W_dev [n_seg][Dx][Dy], bias_dev [n_seg][Dy] or NULL
 * (every held-out trial is predicted by the model that never saw it). */
int td_lagged_predict_per_segment_f32(const td_stream* x, const int64_t* seg_keep,
                                      const int64_t* out_row, int n_seg, const float* W_dev,
                                      const float* bias_dev, int Dy, float* out_dev, int64_t ldo,
                                      void* cuda_stream);

/* The lagged rows themselves (compatibility iterator / tests only; brain_data.py:453-455). */
int td_lagged_rows_f32(const td_stream* x, const int64_t* seg_keep, const int64_t* out_row,
                       int n_seg, float* out_dev, int64_t ldo, void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * Solves (cuSOLVER/cuBLAS inside, float64 internally -- stated non-hot-path dependency).
 *
 * td_ridge_solve: brain_model.py:447-454, :477-479.  A = G_aug / n + lambda I over ALL rows
 * (bias row included), B = C_aug / n, solution = A^-1 B for every lambda.
 *   G_aug_dev [Da, Da] double (Da = D+1 with the ones column last, or D when no offset)
 *   C_aug_dev [Da, Dy] double;  W_out_dev [n_lambda, Da, Dy] float32.
 *   method 0: cuSOLVER Cholesky per lambda on side streams (LU if not positive definite);
 *   1: one eigendecomposition reused for the whole sweep; 2: the whole sweep as ONE batched,
 *   blocked Cholesky (own panel kernels + cublasDgemmStridedBatched; falls back to method 0 for a
 *   lambda whose matrix is not positive definite).  Synchronises the stream.
 * ------------------------------------------------------------------------------------- */
int td_ridge_solve(const double* G_aug_dev, const double* C_aug_dev, int Da, int Dy, double n,
                   const double* lambdas, int n_lambda, int method, float* W_out_dev,
                   void* cuda_stream);

/* td_ridge_solve_batched: n_batch DIFFERENT systems (leave-one-trial-out folds), one lambda each, in
 * one batched blocked Cholesky: G_aug_dev [n_batch][Da][Da], C_aug_dev [n_batch][Da][Dy] double
 * (un-normalised sums), n and lambdas HOST arrays [n_batch]; W_out_dev [n_batch][Da][Dy] float32.
 * A system that is not positive definite is an error.  Synchronises the stream. */
int td_ridge_solve_batched(const double* G_aug_dev, const double* C_aug_dev, int Da, int Dy,
                           const double* n, const double* lambdas, int n_batch, float* W_out_dev,
                           void* cuda_stream);

/* td_ridge_loo_solve: every leave-one-trial-out fold from ONE factorisation.  Fold f solves
 * (G_tot - G_f + lambda n_f I) w = C_tot - C_f; with S = G_tot + lambda n_ref I that is (S - E_f) w = c_f,
 * E_f = G_f + lambda (n_ref - n_f) I small (one trial of F), and w <- S^-1 (c_f + E_f w) contracts by
 * ~1/F per step.  Inputs: G_aug_total_dev [D+1][D+1], C_aug_total_dev [D+1][Dy] double (ones column
 * last), the held-out trials' statistics gram_fold_dev [F][D][D] float32 (td_lagged_gram_batched_f32),
 * xtv_fold_dev [F][D][Dy+1], vtv_fold_dev [F][Dy+1][Dy+1] double (td_lagged_thin_per_segment_f64),
 * n_fold HOST [F] rows per trial, n_total.  W_out_dev [F][D+1][Dy] float32.  *iterations and
 * *worst_rel (HOST, may be NULL) report the steps taken and the largest relative change of the last
 * one; the status is TD_OK even when that is above tol -- the caller may then fall back to
 * td_ridge_solve_batched.  Dy <= 4.  Synchronises the stream. */
int td_ridge_loo_solve(const double* G_aug_total_dev, const double* C_aug_total_dev,
                       const float* gram_fold_dev, const double* xtv_fold_dev,
                       const double* vtv_fold_dev, const double* n_fold, int D, int Dy, int F,
                       double n_total, double lambda, int max_iter, double tol, float* W_out_dev,
                       int* iterations, double* worst_rel, void* cuda_stream);

/* td_cca_solve: cca.py:337-367.  Inputs are raw sums (double): Gxx [D1,D1], Gyy [D2,D2],
 * Gxy [D1,D2], sum_x [D1], sum_y [D2]; n_frames = total rows, n_norm = n_batches*n_row - 1.
 * Outputs float32: rot_x [D1,dim], rot_y [D2,dim], mean_x [D1], mean_y [D2], e [dim].
 * Signs: each (rot_x[:,i], rot_y[:,i]) pair is normalised so that the entry of rot_x[:,i] with
 * the largest magnitude is positive.  With reg > 10*eps_eig the whitening is by Cholesky factors
 * (same canonical pairs, ~10x faster); otherwise, or if a factorisation fails or the dim-th
 * canonical correlation is below 1e-4 of the first, by the reference's eigendecompositions with the
 * eps_eig threshold and an SVD.  Synchronises the stream. */
int td_cca_solve(const double* Gxx_dev, const double* Gyy_dev, const double* Gxy_dev,
                 const double* sum_x_dev, const double* sum_y_dev, int D1, int D2,
                 double n_frames, double n_norm, double reg, double eps_eig, int dim,
                 float* rot_x_dev, float* rot_y_dev, float* mean_x_dev, float* mean_y_dev,
                 float* e_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * Correlation / windows / decision.
 *
 * td_column_moments_f64: the Decoder's streaming statistics, infer_decoder.py:299-303.
 *   sums_dev [4, dims] double += [sum a, sum b, sum a^2, sum b^2] over n rows (NOT zeroed).
 * ------------------------------------------------------------------------------------- */
int td_column_moments_f64(const float* a_dev, int64_t lda, const float* b_dev, int64_t ldb,
                          int64_t n, int dims, double* sums_dev, void* cuda_stream);

/* td_windowed_corr_f32: sliding windows [i*S, i*S+W) per segment while W frames remain
 * (result_store.py:262-271; count (n-W)//S + 1).
 *   mode 0  plain:   value = float64 mean over the window of a[t, 0]           (infer.py:264-265)
 *   mode 1  global:  c[t,d] = (a-mu_a)(b-mu_b)/power in float32 (infer_decoder.py:327-328),
 *                    reduced over d by `reduction` (0 first, 1 second, 2 mean, 3 mean-squared;
 *                    infer_decoder.py:441-450), then float64 window mean
 *                    reduction 5 ('lda'): slope * sum_d c[t,d] w[d] + intercept in float64
 *                    (scaled_lda.py:197-212, :345-355), lda_params_dev = [w (dims), slope, intercept]
 *   mode 2  local:   Pearson r per window and column from the five sums
 *                    (brain_model.py:62-79); out has `dims` columns, and a window whose
 *                    product of powers over columns is <= 0 yields zeros
 *   mode 3  global, every column: window means of c[t,d] for each d ('all' reduction, and
 *                    average_data of the training pass, infer_decoder.py:377-400, :748-783)
 * seg_start/seg_len HOST arrays (rows of a/b); n_sizes window shapes (W[j], S[j]) are evaluated
 * in ONE pass over the data.  out_dev is double [total_windows, out_cols]; windows are ordered
 * size-major, then segment, then window index; out_offset (HOST, n_sizes+1 entries, may be
 * NULL) receives the first window row of each size. */
int td_windowed_corr_f32(const float* a_dev, int64_t lda, const float* b_dev, int64_t ldb,
                         int dims, const int64_t* seg_start, const int64_t* seg_len, int n_seg,
                         const int64_t* W, const int64_t* S, int n_sizes, int mode, int reduction,
                         const float* mu_a_dev, const float* mu_b_dev, const float* power_dev,
                         const double* lda_params_dev, double* out_dev, int64_t* out_offset,
                         void* cuda_stream);
int64_t td_window_count(int64_t n_frames, int64_t W, int64_t S);

/* Winner-take-all: out[i] = r1[i] > r2[i] (attention_decoder.py:134). */
int td_wta_decide_u8(const double* r1_dev, const double* r2_dev, int64_t n, uint8_t* out_dev,
                     void* cuda_stream);

/* Utility: out[j, i] = out[i, j] for j > i (mirror the computed triangle). */
int td_symmetrize_f32(float* G_dev, int64_t n, int64_t ldg, int from_upper, void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, every rank accumulates the statistics of its share of the frames
 * (seg_first / seg_keep above) and ONE allreduce over NVLink combines them.  Replaces nothing in the
 * reference line by line -- it has no multi-device path; the sharded pass stands in for its
 * per-(lambda, fold) re-read of the data (regression.py:381-409).
 * NCCL is bound at run time (dlopen of libnccl.so.2; the copy the process has already loaded if any).
 *   td_comm_unique_id  128 bytes to hand to every rank (rank 0 makes it, the host distributes it)
 *   td_comm_init       collective over `world` ranks; the current CUDA device is the rank's GPU
 *   td_allreduce_sum_* in place, on `cuda_stream`
 *   td_allreduce_sym_f32  symmetric [n, n] matrix: packs the upper triangle (n(n+1)/2 floats) into
 *                      packed_dev, reduces that, unpacks into both triangles -- half the bytes on the wire
 * ------------------------------------------------------------------------------------- */
typedef struct td_comm td_comm;
int td_comm_unique_id(void* id128);
int td_comm_init(td_comm** comm, const void* id128, int rank, int world);
int td_comm_destroy(td_comm* comm);
int td_comm_nccl_version(int* version);
int td_allreduce_sum_f32(td_comm* comm, float* buf_dev, int64_t n, void* cuda_stream);
int td_allreduce_sum_f64(td_comm* comm, double* buf_dev, int64_t n, void* cuda_stream);
int td_allreduce_sym_f32(td_comm* comm, float* G_dev, int64_t n, int64_t ldg, float* packed_dev,
                         size_t packed_bytes, void* cuda_stream);
int td_pack_upper_f32(const float* G_dev, int64_t n, int64_t ldg, float* packed_dev, void* cuda_stream);
int td_unpack_upper_f32(const float* packed_dev, int64_t n, int64_t ldg, float* G_dev, void* cuda_stream);

/* ---------------------------------------------------------------------------------------
 * TFRecord / tf.train.Example reader (HOST code, no GPU needed).  Replaces
 * tf.data.TFRecordDataset + tf.io.parse_example as the reference uses them to read one frame per
 * record (brain_data.py:733-745, :758-812), discover_feature_shapes (:853-892) and count_tfrecords
 * (:895-920); wire format written by ingest.py:1118-1172.  `file_bytes` is the whole file in host
 * memory.  verify_crc: 0 none, 1 the length checksums, 2 length and data checksums (masked CRC32C).
 *   td_tfrecord_count     number of well-framed records (on an error: those before the damage).
 *   td_tfrecord_features  names ('\0'-separated into `names`), value counts and kinds (1 bytes,
 *                         2 float, 3 int64) of the features of the FIRST record.
 *   td_tfrecord_read_f32  out[r, :] = the requested features of record r side by side (widths[i]
 *                         values each, int64 converted), row stride ld_out floats; a record that
 *                         lacks a feature or holds another number of values is an error.
 * ------------------------------------------------------------------------------------- */
int td_tfrecord_count(const void* file_bytes, size_t n_bytes, int verify_crc, int64_t* n_records);
int td_tfrecord_features(const void* file_bytes, size_t n_bytes, char* names, size_t names_cap,
                         int32_t* widths, int32_t* kinds, int32_t max_features, int32_t* n_features);
int td_tfrecord_read_f32(const void* file_bytes, size_t n_bytes, const char* const* names,
                         const int32_t* widths, int32_t n_names, float* out_host, int64_t ld_out,
                         int64_t max_records, int verify_crc, int64_t* n_records);

#ifdef __cplusplus
}
#endif
#endif  /* TDGRAM_H_ */
