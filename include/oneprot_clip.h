/* oneprot_clip.h - C ABI of the B200-native ClipLoss hot path (liboneprot_clip.so).
 *
 * The reference (klemens-floege/oneprot) has no FFI for this path: the boundary is the Python
 * class ClipLoss (src/models/components/loss.py:49-114) plus gather_features (loss.py:19-46) and
 * the Normalize / LearnableLogitScaling epilogue (src/models/components/base_encoder.py:6-33).
 * Each entry point below states which reference lines it replaces.  All pointers are raw device
 * pointers owned by the caller (PyTorch's caching allocator in the Python host); the library
 * allocates no device memory and keeps no state besides a thread-local error string.  Every
 * function returns 0 on success and a non-zero code otherwise (oneprot_last_error() describes
 * it); nothing throws across this boundary.  `stream` is a cudaStream_t passed as void*.
 *
 * Notation: n = rows held by this rank, N = global rows (= world_size * n), d = feature dim,
 * row_offset = rank * n, A = first positional feature tensor (n x d, row-major, bf16),
 * B_all = second positional feature tensor gathered over ranks (N x d, row-major, bf16).
 *   x_ij = c * <a_i, b_j>,  c = logit_scale * log2(e)   (logits in log2 units, fp32 accumulators)
 *   e_ij = 2^(x_ij - G),    G = max(0, |c| * max|a| * max|b| - 100)  (one global reference)
 */
#ifndef ONEPROT_CLIP_H
#define ONEPROT_CLIP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONEPROT_ABI_VERSION 1

/* error codes */
#define ONEPROT_OK 0
#define ONEPROT_ERR_ARG 1      /* bad argument (null pointer, misaligned, unsupported shape) */
#define ONEPROT_ERR_CUDA 2     /* CUDA runtime / driver error */
#define ONEPROT_ERR_DEVICE 3   /* device is not sm_100 (no tcgen05 / TMEM) */

int oneprot_abi_version(void);
const char* oneprot_last_error(void);
/* 0 iff `device` exists and is compute capability 10.x */
int oneprot_device_check(int device);
/* SM count of the current device (persistent grids launch one CTA per SM; the host sizes dL/dZ
 * panels so that the dA GEMM's 128 x 256 tiles fill whole waves). */
int oneprot_num_sms(void);
/* number of kernel launches issued by this library (all threads) since the last reset */
long long oneprot_launch_count(void);
void oneprot_launch_count_reset(void);

/* ---- gradient / value conventions --------------------------------- */
#define ONEPROT_MODE_GLOBAL 0  /* local_loss = False : value = global loss                    */
#define ONEPROT_MODE_LOCAL 1   /* local_loss = True  : value = mean over this rank's rows/cols */

/* ---- forward ------------------------------------------------------------------------------ */

/* Row statistics: diag[i] = <a_i, b_{row_offset+i}> (fp32), stats[0] = max_i |a_i|^2,
 * stats[1] = max_j |b_j|^2 over B_all (atomic max; stats has FOUR floats, all zero-initialised by
 * the caller; [2..3] belong to the max pass of oneprot_clip_fwd_sums).
 * Replaces nothing 1:1 - it supplies the label logits that F.cross_entropy gathers
 * (loss.py:109-112 with labels from get_ground_truth, loss.py:72-83). */
int oneprot_clip_rowstats(const void* A, const void* B_all, int n, int N, int d, int row_offset,
                          float* diag, float* stats, void* stream);

/* Bytes of scratch oneprot_clip_fwd_sums needs for a (n x N) logit panel. */
size_t oneprot_clip_fwd_scratch_bytes(int n, int N);

/* Fused logit GEMM + exp-sum epilogue (tcgen05/TMEM/TMA): for the row panel
 * Z[row_offset : row_offset+n, 0:N] computes, without ever storing a logit,
 *   rowsum[i] = sum_j e_ij (complete for this rank's rows)  and
 *   colsum[j] = sum_{i in panel} e_ij (partial over ranks; sum them across ranks).
 * Replaces get_logits (loss.py:85-101) + the log_softmax half of F.cross_entropy (loss.py:109-112).
 * scale_dev: device pointer to the fp32 logit_scale; stats: FOUR floats, [0..1] as written by
 * rowstats (after a cross-rank max when world_size > 1), [2..3] zero on entry.
 * Robust tier: when n == N (whole matrix on this GPU) and |c| max|a| max|b| > 100 a max pass
 * (same tensor-core mainloop, max epilogue) first writes the exact maximum logit to stats[2] and
 * sets stats[3] = 1; every later kernel then uses G = max(0, stats[2] - 100).  The pass is always
 * enqueued and returns immediately when the bound is rigorous (no host synchronisation). */
int oneprot_clip_fwd_sums(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev,
                          float* stats, float* rowsum, float* colsum, void* scratch, size_t scratch_bytes,
                          void* stream);

/* Same kernel with the all-gather of the second operand FUSED in (single NVSwitch node): the
 * spare warp of every CTA pushes a slice of this rank's rows (src, rows_per_rank x d bf16) through
 * the NVLink multicast alias dst_mc of their place inside B_all, in `chunks` row chunks, each
 * followed by a multicast flag; the TMA producer of every GPU waits for the flag of a column
 * block (flags == epoch) before loading it, so tiles of chunks that have landed are computed while
 * later chunks are still in flight.  The norm maxima (`stats`, local rows only) travel the same
 * way; their global maximum is written to stats_out[0..1] for the later kernels.
 * flags/flags_mc: u32[world][chunks + 1]; stats_all/stats_mc: float[world][4]; counters:
 * u32[chunks], zero-initialised once.  All *_mc pointers alias symmetric memory.
 * The flags of a chunk are raised only after EVERY CTA of the grid has pushed its slice, so the whole grid must be
 * resident at once: with `ag` the kernel is launched cooperatively (the launch fails instead of starting a grid that
 * cannot be placed) and every flag wait is a bounded spin that traps (ONEPROT_WAIT_TRAP_CYCLES).
 * Replaces gather_features (loss.py:19-46) for the forward, overlapped with get_logits. */
typedef struct {
  const void* src;
  void* dst_mc;
  unsigned int* counters;
  unsigned int* flags_mc;
  const unsigned int* flags;
  float* stats_mc;
  const float* stats_all;
  float* stats_out;
  unsigned int epoch;
  int rank, world, chunks, rows_per_rank;
} oneprot_ag_t;
int oneprot_clip_fwd_sums_ag(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev,
                             float* stats, const oneprot_ag_t* ag, float* rowsum, float* colsum, void* scratch,
                             size_t scratch_bytes, void* stream);

/* Loss value + softmax normalisers from complete sums (all length N, global index order):
 *   loss_out[0] = this rank's return value (MODE_GLOBAL: mean over all N; MODE_LOCAL: mean over
 *   rows/cols [row_offset, row_offset+n)), fp32;  inv_rowsum/inv_colsum[k] = 1/sum;
 *   flag[0] |= 1 if any sum left the validated fp32 window; loss_out[0] is then NaN (not trustworthy).
 * scratch: ONEPROT_FINALIZE_SCRATCH_BYTES bytes, 8-byte aligned, zero-initialised ONCE by the
 * caller (the kernel leaves it ready for the next call).
 * Replaces the nll half of F.cross_entropy and the /2 (loss.py:109-112). */
#define ONEPROT_FINALIZE_SCRATCH_BYTES 512
int oneprot_clip_loss_finalize(const float* rowsum_all, const float* colsum_all, const float* diag_all, int N,
                               int n, int row_offset, int mode, const float* scale_dev, const float* stats,
                               float* loss_out, float* inv_rowsum, float* inv_colsum, int* flag, void* scratch,
                               void* stream);

/* ---- two-reference (robust) path: inputs whose row / column maxima are too far apart for one
 * common reference.  (1) oneprot_clip_rowcol_max: per-row maxima of the panel (complete) and
 * per-column maxima over this rank's rows (max them across ranks), in log2 units, same
 * tensor-core mainloop as the forward.  (2) oneprot_augment_bf16 appends 8 columns to an operand:
 * [x | e_h | e_m | 0..0] with e_h + e_m = -ref/c in two bf16 limbs, or e_h = e_m = 1, so that a GEMM
 * over d + 8 columns yields x_ij - ref'_i with ref'_i = -c * (float(e_h) + float(e_m)) (returned in
 * ref_q; within 2^-17 |ref| of ref).  The forward / dz / GEMM kernels
 * then run unchanged on the augmented operands with G = 0 (stats = [*, *, 0, 1]), once per softmax
 * direction.  (3) oneprot_clip_loss_finalize_ex adds the per-element references back:
 * LSE_row_i = ln2 * (row_ref[i] + log2 rowsum_i). */
int oneprot_clip_rowcol_max(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev, float* rowmax,
                            float* colmax, void* scratch, size_t scratch_bytes, void* stream);
int oneprot_augment_bf16(const void* in, int rows, int d, const float* ref, const float* scale_dev, void* out, float* ref_q,
                         void* stream);
int oneprot_clip_loss_finalize_ex(const float* rowsum_all, const float* colsum_all, const float* diag_all, int N, int n,
                                  int row_offset, int mode, const float* scale_dev, const float* stats, float* loss_out,
                                  float* inv_rowsum, float* inv_colsum, int* flag, void* scratch, const float* row_ref,
                                  const float* col_ref, void* stream);

/* ---- retrieval metric (src/models/components/retrieval_metric.py:76-102) ----
 * rank_s2m[i] = #{ j != i : <s_i, m_j> > <s_i, m_i> }  (position of the label in the descending
 * argsort of logits_per_sequence row i when there are no ties), rank_m2s[j] likewise over the
 * columns - from the logits tiles of the same tensor-core mainloop, never materialising S M^T.
 * label_dot[i] = <s_i, m_i> as written by oneprot_clip_rowstats (diag).  S, M: N x d bf16. */
int oneprot_retrieval_ranks(const void* S, const void* M, int N, int d, const float* label_dot, float* rank_s2m,
                            float* rank_m2s, void* scratch, size_t scratch_bytes, void* stream);

/* ---- SigLipLoss (loss.py:204-311), the other objective behind OneProt's loss_fn switch
 * (oneprot_module.py:57-62):  value_r = -(1/n) sum_{i in rank r, j} logsigmoid(y_ij z_ij),
 * z_ij = logit_scale <a_i, b_j> + logit_bias, y = +1 on the global diagonal and -1 elsewhere.
 * Since -logsigmoid(-z) = softplus(z) and -logsigmoid(z) = softplus(z) - z, the tile epilogue only
 * sums softplus(z) and the label term comes from the diag of oneprot_clip_rowstats:
 *   oneprot_siglip_fwd       rowsum[i] = sum_j log2(1 + 2^x_ij), x = log2(e) z (same tensor-core mainloop
 *                            as the ClipLoss forward; scratch of oneprot_clip_fwd_scratch_bytes(n, N))
 *   oneprot_siglip_finalize  loss_out[0] = (ln2 sum_i rowsum[i] - sum_i (logit_scale diag[i] + bias)) / n
 *   oneprot_siglip_dz_panel  Wz_ij = wr[i] sigma(z_ij) - [grow0 + i == j] dg[i]  (bf16 panel, as
 *                            oneprot_clip_dz_panel; dL/dz_ij = (g / n) (sigma(z_ij) - [i == j])).
 *                            sig_rowsum (optional, with scratch of oneprot_siglip_dz_scratch_bytes):
 *                            sig_rowsum[i] = sum_j sigma(z_ij), for d logit_bias = (g / n) (sum sigma - n)
 * The ring of neighbour exchanges of the reference (loss.py:258-309) becomes the all-gather of the
 * second operand + reduce-scatter of its partial gradient that the ClipLoss path already uses.
 * bias_dev may be NULL (no logit_bias). */
int oneprot_siglip_fwd(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev, const float* bias_dev,
                       float* rowsum, void* scratch, size_t scratch_bytes, void* stream);
/* Kept-panel variant (opt-in): also writes S[i][j] = sigma(z_ij) - [grow0 + i == j] as bf16 (n rows, row pitch lds >= N,
 * multiple of 8) - dL/dz up to the constant g / n, so the backward is the two GEMMs on S as it is - and, when sig_rowsum
 * is not NULL, sig_rowsum[i] = sum_j sigma(z_ij).  Scratch: oneprot_siglip_fwd_keep_scratch_bytes(n, N). */
size_t oneprot_siglip_fwd_keep_scratch_bytes(int n, int N);
int oneprot_siglip_fwd_keep(const void* A, const void* B_all, int n, int N, int d, int grow0, const float* scale_dev,
                            const float* bias_dev, float* rowsum, float* sig_rowsum, void* scratch, size_t scratch_bytes,
                            void* S, int lds, void* stream);
int oneprot_siglip_finalize(const float* rowsum, const float* diag, int n, const float* scale_dev, const float* bias_dev,
                            float* loss_out, void* stream);
size_t oneprot_siglip_dz_scratch_bytes(int rows, int N);
int oneprot_siglip_dz_panel(const void* A_rows, const void* B_all, int rows, int N, int d, int grow0, const float* scale_dev,
                            const float* bias_dev, const float* wr, const float* dg, void* Wz, int ldw, float* sig_rowsum,
                            void* scratch, size_t scratch_bytes, void* stream);

/* ---- backward ----------------------------------------------------------------------------- */

/* Per-row / per-column / diagonal coefficients of dL/dZ for the panel of this rank:
 *   Wz_ij = e_ij * (wr[i] + wc[j]) - [i == j] * dg[i]      (already multiplied by logit_scale)
 * gvec_dev[world] holds the upstream gradient of every rank (device fp32; world = 1: one value).
 * use_gsum != 0 applies the reduce-scatter-SUM convention of torch.distributed.nn.all_gather's
 * backward (loss.py:32-33): coefficients carry sum_r g_r (MODE_GLOBAL) or g_owner (MODE_LOCAL).
 * part: 0 = both softmax directions, 1 = row-softmax part only, 2 = column-softmax part only
 * (the two halves of the local_loss=True, gather_with_grad=False convention).
 * MODE_GLOBAL writes a unit-gradient panel description and the upstream gradients go to
 * out_scale_a[n] (multiplies the rows of dA) and out_scale_b[N] (multiplies the rows of the
 * partial dB): g_own resp. g_owner(j) without, sum_r g_r with use_gsum.  MODE_LOCAL folds the
 * gradients into wr/wc and writes ones.
 * what: 0 = write everything, 1 = only wr/wc/dg, 2 = only the output scales (lets the host
 * overlap the exchange of the upstream gradients with the dL/dZ panel kernel). */
int oneprot_clip_bwd_weights(const float* inv_rowsum, const float* inv_colsum, int N, int n, int row_offset,
                             int mode, int use_gsum, int part, int world, int rank, const float* gvec_dev,
                             const float* scale_dev, float* wr, float* wc, float* dg, float* out_scale_a,
                             float* out_scale_b, int what, void* stream);

/* Recompute logit tiles for rows [r0, r0+rows) of this rank's panel and write
 * Wz (bf16, row-major, leading dimension ldw >= N, multiple of 8) - the bounded dL/dZ panel.
 * A_rows points at row r0 of A; wr/dg point at element r0.  grow0 = row_offset + r0 (global row
 * of the first panel row, for the diagonal).  Replaces the autograd backward of
 * F.cross_entropy (loss.py:109-112).
 * PADDING CONTRACT (every bf16 panel written through a TMA store: Wz here, E / S of the *_keep entry points):
 * rows beyond `rows` are never touched; the padding columns [N, ldw) of the written rows are SCRATCH - when N is not
 * a multiple of 8 elements the hardware store does not clip the last 16-byte chunk per element (measured on B200,
 * round 1) - and no entry point of this library reads them (every consumer's tensor map has inner extent N). */
int oneprot_clip_dz_panel(const void* A_rows, const void* B_all, int rows, int N, int d, int grow0,
                          const float* scale_dev, const float* stats, const float* wr, const float* wc,
                          const float* dg, void* Wz, int ldw, void* stream);

/* Stored-exponentials variant of the pair above (opt-in; trades n x lde x 2 bytes of HBM for the second
 * pass over the logits): oneprot_clip_fwd_sums_keep is oneprot_clip_fwd_sums_ag (ag may be NULL) that also
 * writes E[i][j] = 2^(x_ij - G) as bf16 (n rows, row pitch lde >= N, multiple of 8, 16-byte aligned; E == NULL:
 * plain forward); oneprot_clip_dz_from_exp then turns rows [r0, r0 + rows) of E into the dL/dZ panel IN PLACE,
 * Wz_ij = E_ij (wr[i] + wc[j]) - [grow0 + i == j] dg[i] (E points at row r0, wr / dg at element r0) - the
 * same panel oneprot_clip_dz_panel writes, up to one more bf16 rounding of e_ij.  HBM-bound, no tensor cores. */
int oneprot_clip_fwd_sums_keep(const void* A, const void* B_all, int n, int N, int d, const float* scale_dev,
                               float* stats, const oneprot_ag_t* ag, float* rowsum, float* colsum, void* scratch,
                               size_t scratch_bytes, void* E, int lde, void* stream);
int oneprot_clip_dz_from_exp(void* E, int rows, int N, int lde, int grow0, const float* wr, const float* wc,
                             const float* dg, void* stream);

/* C[M x Nc] = op(A) * op(B) with bf16 operands, fp32 accumulation in TMEM.
 *   a_mn = 0: A is M x K row-major (lda >= K);  a_mn = 1: A is K x M row-major (lda >= M)
 *   b_mn = 0: B is Nc x K row-major (ldb >= K); b_mn = 1: B is K x Nc row-major (ldb >= Nc)
 * value = acc + (acc_in ? acc_in[m*ldc+n] : 0); stored to acc_out (fp32) and/or out_bf16 (bf16),
 * whichever is non-null, both with leading dimension ldc.  Replaces the autograd backward of
 * the logits matmul (loss.py:92-99): dA = Wz * B_all (a_mn=0,b_mn=1), dB = Wz^T * A (a_mn=1,b_mn=1). */
int oneprot_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int Nc, int K,
                      const float* acc_in, float* acc_out, void* out_bf16, int ldc, void* stream);

/* Same GEMM with two optional epilogue extras:
 *   row_scale[M]  : stored value = row_scale[m] * (acc + acc_in)
 *   dot_mat (bf16, M x Nc, leading dimension ld_dot) + rowdot_part: per-row dot products of the
 *   UNSCALED value with dot_mat, one partial per 128-column slab:
 *   rowdot_part[(slab) * ldd + m], ldd = ceil(M/128)*128, slabs = 2*ceil(Nc/256)
 *   (oneprot_gemm_rowdot_scratch_bytes gives the size).  Used for d logit_scale =
 *   (1/scale) * sum_i <a_i, dA_i>  (backward of the `logit_scale *` in loss.py:92-99). */
size_t oneprot_gemm_rowdot_scratch_bytes(int M, int Nc);
int oneprot_gemm_bf16_ex(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int Nc, int K,
                         const float* acc_in, float* acc_out, void* out_bf16, int ldc, const float* row_scale,
                         const void* dot_mat, int ld_dot, float* rowdot_part, void* stream);

/* out[i] = <x_i, y_i> over d bf16 elements (leading dimensions ldx, ldy; multiples of 8). */
int oneprot_rowdot_bf16(const void* x, int ldx, const void* y, int ldy, int rows, int d, float* out, void* stream);
/* out[0] = sum of v[0..count) in a fixed order (deterministic). */
int oneprot_sum_f32(const float* v, int count, float* out, void* stream);

/* ---- L2-normalise / logit-scale epilogue (base_encoder.py:6-33) -------------------------- */

/* y = scale * x / max(|x|_2, eps) row-wise; x, y: rows x d bf16 (or fp32 when is_fp32);
 * inv_norm[rows] (fp32) is saved for the backward.  scale_dev may be NULL (scale = 1).
 * Replaces Normalize.forward (base_encoder.py:11-12) fused with
 * LearnableLogitScaling.forward (base_encoder.py:29-30; the clip(exp(log_s), max) is computed by
 * the caller into scale_dev). */
int oneprot_l2norm_scale_fwd(const void* x, void* y, float* inv_norm, int rows, int d, int is_fp32,
                             const float* scale_dev, float eps, void* stream);
/* gx = scale * inv_norm * (gy - yhat * <yhat, gy>), yhat = x * inv_norm;
 * dscale_partial[row] = <yhat, gy> (sum over rows = d loss / d scale), may be NULL. */
int oneprot_l2norm_scale_bwd(const void* x, const void* gy, const float* inv_norm, void* gx, float* dscale_partial,
                             int rows, int d, int is_fp32, const float* scale_dev, float eps, void* stream);

/* y = scale * x over rows x d elements (bf16, or fp32 when is_fp32).  Replaces the multiply of
 * LearnableLogitScaling.forward (base_encoder.py:29-30) when it is used without Normalize. */
int oneprot_scale_rows(const void* x, void* y, int rows, int d, int is_fp32, const float* scale_dev, void* stream);
/* out[row] = <x_row, y_row> for contiguous rows (bf16 or fp32): d scale = sum_rows <x, gy>. */
int oneprot_rowdot(const void* x, const void* y, int rows, int d, int is_fp32, float* out, void* stream);

/* ---- single-node exchanges through an NVLink multicast (NVLS) mapping -----------------------
 * The *_mc pointers are multicast aliases of a symmetric buffer (the Python host obtains them
 * from torch.distributed._symmetric_memory); they replace the collectives of gather_features
 * (loss.py:32-38) and of its autograd backward on one NVSwitch node.  The caller brackets them
 * with the symmetric-memory barrier. */
/* all-gather by push: copy `bytes` from src into every GPU's copy of the buffer behind dst_mc */
int oneprot_mc_store(const void* src, void* dst_mc, size_t bytes, void* stream);
/* dst[i] = reduce over GPUs of src_mc[i]; op 0 = fp32 add, 1 = max of non-negative fp32 */
int oneprot_mc_allreduce_f32(const float* src_mc, float* dst, int count, int op, void* stream);
/* dst = sum over GPUs of a bf16 buffer (fp32 accumulation in the switch); reduce-scatter by pull */
int oneprot_mc_reduce_bf16(const void* src_mc, void* dst, size_t bytes, void* stream);

/* Split fp32 rows into bf16 limbs for the fp32-accurate path: out is rows x (terms*d) bf16 with
 * the limb order given by `pattern`, so that the bf16 GEMM over the
 * concatenated K reproduces the fp32 dot product. */
int oneprot_split_fp32(const float* x, void* out, int rows, int d, int side, int terms, void* stream);

/* ---- projection heads in front of the path (base_encoder.py:107-194) --------
 * HBM-bound row kernels (oneprot_b200/csrc/head_kernels.cu); x / y / gradients are bf16, or fp32 when
 * is_fp32, and gamma / beta have the dtype of x.  The Linear layers run on oneprot_gemm_bf16_ex. */

/* torch.nn.LayerNorm over the last dim (base_encoder.py:148,154,157): y = (x - mean) * rstd * gamma + beta
 * with the biased variance and eps inside the root; mean / rstd (fp32, one per row) are saved for
 * the backward.  d: multiple of 8 (rows of up to 2048 elements stay in registers; longer rows -
 * ESM-2 3B / 15B: 2560 / 5120 - are re-read from L1 / L2). */
int oneprot_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd, int rows, int d,
                          int is_fp32, float eps, void* stream);
/* gx = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = gy * gamma (gx may be NULL);
 * dgamma = sum_rows gy * xhat, dbeta = sum_rows gy (fp32, both or neither; summed over
 * per-row-chunk partials in `scratch` in a fixed order). */
size_t oneprot_layernorm_bwd_scratch_bytes(int rows, int d);
int oneprot_layernorm_bwd(const void* x, const void* gy, const void* gamma, const float* mean, const float* rstd, void* gx,
                          float* dgamma, float* dbeta, void* scratch, size_t scratch_bytes, int rows, int d, int is_fp32,
                          void* stream);
/* exact (erf) GELU of base_encoder.py:156: gy == NULL: out = gelu(x); else out = gy * gelu'(x).  count % 8 == 0. */
int oneprot_gelu(const void* x, const void* gy, void* out, size_t count, int is_fp32, void* stream);
/* MeanPooling.forward (base_encoder.py:107-118): y[b] = sum_l mask[b,l] x[b,l,:] / sum_l mask[b,l]
 * for x of shape B x L x D; mask fp32 B x L or NULL (plain mean); inv_count[b] = 1 / sum_l mask[b,l]. */
/* normalize = 0: plain weighted sum y[b] = sum_l mask[b,l] x[b,l,:] (attention pooling; inv_count may be NULL) */
int oneprot_meanpool_fwd(const void* x, const float* mask, void* y, float* inv_count, int B, int L, int D, int is_fp32,
                         int normalize, void* stream);
/* gx[b,l,:] = mask[b,l] * inv_count[b] * gy[b,:] */
int oneprot_meanpool_bwd(const void* gy, const float* mask, const float* inv_count, void* gx, int B, int L, int D, int is_fp32,
                         void* stream);

/* Attention1dPooling (base_encoder.py:84-104): score[b,l] = <w, x[b,l,:]> + bias (a 1x1 conv to one
 * channel), -inf where mask == 0, softmax over the tokens, weighted sum (oneprot_meanpool_fwd with
 * the probabilities as weights and normalize = 0).
 * oneprot_token_dot: out[b,l] = <vec, x[b,l,:]> + bias[0] (bias may be NULL); vec_per_batch = 0: one
 * vector of D elements (the score layer), 1: B vectors (d p = <g[b], x[b,l]> in the backward);
 * mask (fp32, may be NULL): out = -inf where mask == 0.  vec has the dtype of x. */
int oneprot_token_dot(const void* x, const void* vec, int vec_per_batch, const float* bias, const float* mask, float* out, int B,
                      int L, int D, int is_fp32, void* stream);
/* p[b,:] = softmax(s[b,:]) over L (fp32; in place allowed);  ds = p * (dp - sum_k p_k dp_k) */
int oneprot_softmax_rows(const float* s, float* p, int B, int L, void* stream);
int oneprot_softmax_rows_bwd(const float* p, const float* dp, float* ds, int B, int L, void* stream);
/* gx[b,l,:] = p[b,l] * g[b,:] + ds[b,l] * w[:]  (gradient w.r.t. the tokens) */
int oneprot_attnpool_bwd_x(const void* g, const float* p, const float* ds, const void* w, void* gx, int B, int L, int D, int is_fp32,
                           void* stream);
/* out[k] = sum_s part[s * ld + k], k < count, in slot order (deterministic) */
int oneprot_sum_slots_f32(const float* part, int slots, int ld, int count, float* out, void* stream);
/* L1 regulariser of the training step (oneprot_module.py:43-44, 99-101): out[0] = sum |x| / true_count over `count`
 * elements (multiple of 8; elements beyond true_count must be zero padding), deterministic;
 * backward gx = g[0] sign(x) / true_count. */
size_t oneprot_abs_mean_scratch_bytes(size_t count);
int oneprot_abs_mean_fwd(const void* x, size_t count, size_t true_count, int is_fp32, float* out, void* scratch,
                         size_t scratch_bytes, void* stream);
int oneprot_abs_mean_bwd(const void* x, const float* g, size_t count, size_t true_count, int is_fp32, void* gx, void* stream);

/* ---- host-side step sequencer (oneprot_b200/csrc/clip_sequence.cu) -------------------------------
 * One call enqueues a whole PHASE of ClipLoss.forward / its autograd backward (loss.py:103-114 and
 * the implicit backward): the memsets, the kernels above and the event records /
 * waits between the compute stream and the exchange stream, out of one caller-provided workspace.
 * Same launches, order and streams as the Python host (oneprot_b200/clip_loss.py); a phase ends
 * where the caller has to run a symmetric-memory barrier.  Scope: bf16 operands (d % 8 == 0), one
 * pass over both softmax directions (world 1; local_loss = False; local_loss = True with
 * gather_with_grad = True), logit_scale without gradient, exchanges through the NVLS provider with
 * the all-gather fused into the forward kernel.  Everything else stays on the Python path. */

/* `saved` (float*, 16-byte aligned, ONEPROT_SAVED_HEADER_FLOATS + 2 N floats) is what the backward
 * needs from the forward: [loss | maxima (stats, 4 floats) | hazard flag (int32) | 1/rowsum N | 1/colsum N] */
#define ONEPROT_SAVED_LOSS_AT 0
#define ONEPROT_SAVED_STATS_AT 4
#define ONEPROT_SAVED_FLAG_AT 8
#define ONEPROT_SAVED_HEADER_FLOATS 16

typedef struct {
  const void* A;            /* n x d bf16: this rank's first operand */
  const void* B_all;        /* N x d bf16: second operand of all ranks (world 1: the caller's B) */
  const void* stats_rows;   /* rows oneprot_clip_rowstats reads as its B_all: B_all (world 1) or this rank's B */
  const float* scale;       /* device fp32 logit_scale */
  float* stats;             /* 4 floats written by rowstats: saved + ONEPROT_SAVED_STATS_AT (world 1) or
                               this rank's maxima inside the symmetric buffer (fused gather) */
  float* saved;
  void* zero_ptr;           /* fused gather: symmetric [g | maxima | sums] buffer to clear first; else NULL */
  size_t zero_bytes;
  float* sums;              /* fused gather: this rank's partial [colsum | rowsum | diag] (3 N floats) in
                               symmetric memory; NULL: the sums are complete locally (workspace) */
  const float* sums_mc;     /* multicast alias of `sums`, or NULL */
  const oneprot_ag_t* ag;   /* fused all-gather descriptor (stats_out = saved + ONEPROT_SAVED_STATS_AT), or NULL */
  void* ws;                 /* 16-byte aligned, oneprot_seq_fwd_ws_bytes(n, N) bytes */
  size_t ws_bytes;
  void* stream;
  int n, N, d, row_offset, mode;
  int stats_rows_n;         /* rows of stats_rows */
  int stats_off;            /* row of stats_rows that holds the label of A's row 0 */
  void* E;                  /* optional: keep the exponentials (oneprot_clip_fwd_sums_keep), else NULL */
  int lde;                  /* row pitch of E: ceil(N / 64) * 64 */
} oneprot_fwd_seq_t;

size_t oneprot_seq_fwd_ws_bytes(int n, int N);
/* memsets + rowstats + fused forward (what comes before NvlsComm.complete_sums' barrier) */
int oneprot_seq_fwd_begin(const oneprot_fwd_seq_t* f);
/* [switch-side sum of the partial sums] + loss_finalize */
int oneprot_seq_fwd_end(const oneprot_fwd_seq_t* f);
/* both phases back to back (world 1) */
int oneprot_seq_fwd(const oneprot_fwd_seq_t* f);

typedef struct {
  const void* A;
  const void* B_all;
  const float* scale;
  const float* stats;        /* saved + ONEPROT_SAVED_STATS_AT */
  const float* inv_rowsum;   /* saved + ONEPROT_SAVED_HEADER_FLOATS */
  const float* inv_colsum;   /* ... + N */
  const float* g;            /* device fp32: this rank's upstream gradient */
  void* dA;                  /* n x d bf16 (want_a) */
  void* dB;                  /* N x d bf16: dB (world 1) or the partial-dB region in symmetric memory */
  void* ws;                  /* 16-byte aligned, oneprot_seq_bwd_ws_bytes(...) bytes */
  size_t ws_bytes;
  size_t panel_bytes;        /* bound of the bf16 dL/dZ panel */
  void* stream;              /* compute stream */
  /* exchange through the NVLS provider (world > 1); NULL / 0 otherwise */
  void* side_stream;
  float* g_slot;             /* this GPU's copy of the gradient slot ((world+3)/4*4 floats) in symmetric memory */
  const float* g_slot_mc;    /* its multicast alias */
  const void* dB_mc_mine;    /* multicast alias of this rank's n x d rows of dB */
  void* dB_out;              /* n x d bf16: this rank's rows of the summed dB */
  void* seq;                 /* oneprot_seq_create handle (events) */
  int n, N, d, row_offset, mode, use_gsum, world, rank;
  int want_a, want_b;
  int g_on_side;             /* gather the upstream gradients on the side stream under the dL/dZ kernel (MODE_GLOBAL) */
  void* E;                   /* optional: the exponentials kept by the forward; rescaled in place (oneprot_clip_dz_from_exp)
                                as ONE panel of n rows instead of the panel-wise recompute, else NULL */
  int lde;
} oneprot_bwd_seq_t;

int oneprot_seq_create(void** out);
void oneprot_seq_destroy(void* seq);
size_t oneprot_seq_bwd_ws_bytes(int n, int N, int d, int world, int want_b, size_t panel_bytes);
/* kept_panel != 0: the panel is the caller's E (one panel of n rows), the workspace holds only the vectors */
size_t oneprot_seq_bwd_ws_bytes_ex(int n, int N, int d, int world, int want_b, size_t panel_bytes, int kept_panel);
/* number of dL/dZ panels for a panel_bytes bound (+ rows per full panel, allocated panel rows) */
int oneprot_seq_bwd_panels(int n, int N, int d, size_t panel_bytes, int* rows_per_panel, int* wz_rows);
/* world > 1: this rank's one-hot upstream gradient into the symmetric slot (before the caller's barrier) */
int oneprot_seq_bwd_begin(const oneprot_bwd_seq_t* q);
/* [gather of the upstream gradients] + bwd_weights + per panel: dz_panel, dB GEMM, dA GEMM
 * (world > 1: ends with the side stream waiting for the last dB GEMM, before the caller's barrier;
 * the last panel's dA GEMM is left to oneprot_seq_bwd_end) */
int oneprot_seq_bwd_main(const oneprot_bwd_seq_t* q);
/* world > 1: switch-side reduce of this rank's dB rows on the side stream, the last panel's dA GEMM
 * on the compute stream over it, then the compute stream waits for the reduce */
int oneprot_seq_bwd_end(const oneprot_bwd_seq_t* q);

/* ---- launch trace (test support) -------------------------------------------------------------
 * Between oneprot_trace_begin and oneprot_trace_end every entry point of this library appends one
 * text line with its arguments; with dry_run != 0 it then returns without any CUDA call, so the
 * launch sequence of a phase can be inspected on a machine without a GPU.  oneprot_trace_end
 * copies the text (NUL-terminated, truncated to cap) and returns its full length. */
void oneprot_trace_begin(int dry_run);
size_t oneprot_trace_end(char* out, size_t cap);
void oneprot_trace_note(const char* text);

#ifdef __cplusplus
}
#endif
#endif /* ONEPROT_CLIP_H */
