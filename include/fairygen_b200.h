/* fairygen_b200 — C ABI of the B200-native Wan2.2-TI2V-5B DiT denoise hot path.
 *
 * This is the drop-in boundary: every entry point replaces one group of library-kernel call sites
 * of the reference (CloudEngineHub/FairyGen, animation/diffsynth). The reference has no native
 * code and no FFI of its own — its "operator API" for this path is the set of torch calls made by
 * diffsynth/models/wan_video_dit.py (DIT), diffsynth/pipelines/wan_video.py (PIPE) and
 * diffsynth/diffusion/flow_match.py (FM); each declaration below cites the lines it replaces.
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (PyTorch); the library never allocates
 *    or frees device memory. Matrices are row-major with the last dimension contiguous; `ld*` are
 *    leading dimensions in ELEMENTS. bf16 unless stated. Pointers and leading dimensions of
 *    matrices that feed tensor-core kernels must be 16-byte aligned.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Every call only
 *    enqueues work; nothing synchronises except fgb_sync_check.
 *  - Every function returns 0 on success, a non-zero fgb_status otherwise; fgb_last_error() gives
 *    the message for the calling thread. No exit()/abort() inside the library.
 *  - There is NO CPU fallback: without a CUDA device fgb_create fails.
 */
#ifndef FAIRYGEN_B200_H_
#define FAIRYGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGB_ABI_VERSION 1
#define FGB_HEAD_DIM 128 /* Wan2.2-TI2V-5B: dim 3072 / 24 heads (model_configs.py:294) */

typedef enum fgb_status {
  FGB_OK = 0,
  FGB_ERR_INVALID = 1,     /* bad argument (shape, alignment, null pointer) */
  FGB_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed            */
  FGB_ERR_UNSUPPORTED = 3, /* not an sm_100 device, or unsupported shape     */
} fgb_status;

/* Epilogues of fgb_gemm_bf16. acc = A·Wᵀ in fp32; y = bf16(acc + bias). */
typedef enum fgb_epilogue {
  FGB_EPI_BIAS = 0,           /* C = y                                   nn.Linear, DIT:140-142,176-178 */
  FGB_EPI_BIAS_GELU_TANH = 1, /* C = gelu_tanh(y)                        ffn[0]+ffn[1], DIT:208-209     */
  FGB_EPI_GATED_RESIDUAL = 2, /* C = C + gate[row] * y                   GateModule, DIT:192-193,225,228 */
  FGB_EPI_RESIDUAL = 3,       /* C = C + y                               cross-attn residual, DIT:226   */
} fgb_epilogue;

typedef struct fgb_ctx fgb_ctx;

int fgb_abi_version(void);
const char* fgb_last_error(void);

/* One context per process/GPU (the reference is one process per GPU, xdit_context_parallel.py:21). */
int fgb_create(int device, fgb_ctx** out);
int fgb_destroy(fgb_ctx* ctx);
/* cudaStreamSynchronize + sticky-error check (asynchronous kernel faults surface here). */
int fgb_sync_check(fgb_ctx* ctx, void* stream);
int fgb_sm_count(fgb_ctx* ctx);

/* ---- tensor-core kernels (tcgen05 + TMEM + TMA) ------------------------------------------- */

/* C[m,n] = epilogue(A[m,k] · W[n,k]ᵀ + bias[n]).  W is an nn.Linear weight as stored ([out,in]).
 * Replaces every nn.Linear on the path: q/k/v/o (DIT:140-146, 176-185), ffn (DIT:208-209),
 * head (DIT:258,264), patch_embedding as a GEMM (DIT:305), time/text embeddings (DIT:307-318).
 * GATED_RESIDUAL: row r uses gate0 when r < rows_gate0, else gate1 (the TI2V per-token timestep
 * of PIPE:1218-1228 has only two distinct modulation rows); gate0/gate1 are [n] bf16.
 * bias may be NULL. k*2 bytes and all leading dimensions*2 bytes must be multiples of 16. */
int fgb_gemm_bf16(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias,
                  void* c, int64_t ldc, int32_t m, int32_t n, int32_t k, int32_t epilogue,
                  const void* gate0, const void* gate1, int32_t rows_gate0, void* stream);

/* fgb_gemm_bf16 with a stream-K tail (the DiT block's Linears, DIT:140-146, 176-185, 208-209, at sizes where the tile count is
 * not a multiple of the 74 CTA pairs: e.g. 27 280 x 3072 = 1284 tiles = 17.35 waves, or 6820 x 3072 on a Ulysses SP4 rank =
 * 4.38 waves). The K range of the tiles of the last, partly filled wave is spread over all CTA pairs; partial fp32 accumulators
 * meet in `workspace` (caller-owned, 16-byte aligned, fgb_gemm_workspace_bytes() bytes, ZERO on first use — the kernel leaves its
 * flags zero again). Launches that share a workspace must be ordered on one stream. NULL workspace = fgb_gemm_bf16. The result
 * of a split tile differs from the unsplit one only in fp32 summation order (deterministic for a given shape). */
int fgb_gemm_bf16_sk(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
                     int32_t m, int32_t n, int32_t k, int32_t epilogue, const void* gate0, const void* gate1, int32_t rows_gate0,
                     void* workspace, int64_t workspace_bytes, void* stream);
int64_t fgb_gemm_workspace_bytes(fgb_ctx* ctx);
/* When fgb_gemm_bf16_sk splits: only if k >= min_k (default 6144: the fix-up costs ~15 us, a short-K tile is ~20 us) and the
 * tail fills at most max_tail_frac of the CTA pairs (default 0.9). Tests set min_k = 0 to drive the split path on small shapes. */
int fgb_gemm_streamk_tune(fgb_ctx* ctx, int32_t min_k, double max_tail_frac);
/* Host-only self-check of the work list a 2-CTA GEMM launch of this shape would walk on a GPU with sm_count SMs (no device
 * needed): every (tile, K-block, 128-column half) exactly once, one owner per split tile whose wait list equals the clusters
 * that dump a partial, split tiles last in their owner's list. with_workspace / min_k as fgb_gemm_bf16_sk / _streamk_tune.
 * Returns FGB_OK or FGB_ERR_INVALID with fgb_last_error(); the counts (optional) say how many tiles were split along K and how
 * many half-width tail items there are. Used by the CPU test suite (tests/test_gemm_schedule.py). */
int fgb_gemm_schedule_check(int32_t m, int32_t n, int32_t k, int32_t sm_count, int32_t with_workspace, int32_t min_k,
                            int32_t* n_split_tiles, int32_t* n_half_items);

/* Same with a second operand pair folded in as extra K-blocks:  acc = A·Wᵀ + A2·W2ᵀ  (a2 [m, k2], w2 [n, k2], k2 % 8 == 0).
 * Used by the stage-2 LoRA forward (training_module.py:317-352): y = W₁x + b + (B2*mask*2)(A1 x) without re-merging the
 * weight every step — A2 = A1·x (rank r), W2 = the masked B2. */
int fgb_gemm_bf16_ex(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c,
                     int64_t ldc, int32_t m, int32_t n, int32_t k, int32_t epilogue, const void* gate0, const void* gate1,
                     int32_t rows_gate0, const void* a2, int64_t lda2, const void* w2, int64_t ldw2, int32_t k2, void* stream);

/* o = softmax(q kᵀ · scale) v per head, no mask, non-causal; head_dim = 128.
 * q,o: [s_q, heads*128]; k,v: [s_kv, heads*128] (row strides ldq/ldk/ldv/ldo elements).
 * Replaces flash_attention()/AttentionModule for self- and cross-attention (DIT:27-60, 113-120,
 * 145, 179). Keys at index >= s_kv do not exist (no padding attends). */
int fgb_attn_fwd(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                 int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                 void* stream);

/* Same, with the two extras the multi-GPU and training paths need:
 *  - workspace (optional, caller-owned, 16-byte aligned, fgb_attn_workspace_bytes() bytes): lets the kernel cut the
 *    units of the last, partly filled wave of CTAs into key chunks whose partial (max, sum, O) results are merged by a
 *    small combine kernel. With heads/P heads per rank under Ulysses SP the grid is only 2.2 waves of the 148 SMs at
 *    P = 8, so without the split a third of the machine idles for a whole wave. NULL = no split.
 *  - lse (optional): fp32 [heads][ld_lse], log2-domain log-sum-exp of the scaled scores, saved for fgb_attn_bwd
 *    (ld_lse >= s_q; rows in [s_q, ld_lse) receive the value of an all-zero query row). */
int fgb_attn_fwd_ex(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                    int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                    void* lse, int64_t ld_lse, void* workspace, int64_t workspace_bytes, void* stream);
int64_t fgb_attn_workspace_bytes(fgb_ctx* ctx, int32_t s_q, int32_t s_kv, int32_t heads);
/* Host-only self-check of the work list an attention launch of this shape walks on sm_count SMs (no device needed): every
 * (256-query unit, KV tile) exactly once, non-empty key chunks, partial slots inside the workspace. The counts (optional) say
 * how the last wave is cut: `split` key chunks for each of `n_split_units` units. Used by tests/test_gemm_schedule.py. */
int fgb_attn_schedule_check(int32_t s_q, int32_t s_kv, int32_t heads, int32_t sm_count, int32_t with_workspace, int32_t* split,
                            int32_t* n_split_units);

/* Bounded-score softmax. kmax2[head] = max_j ||k[j, head]||^2 (fgb_head_norm_max) gives |q_i·k_j|·scale·log2e <= B_i =
 * ||q_i||·sqrt(kmax2)·scale·log2e, from which every query row gets a FIXED reference R_i for its exponentials P = 2^(s - R_i):
 * no running maximum, no rescale of O, no per-tile exchange between the threads that share a row. It is the same softmax as
 * long as the row's largest score stays inside the exponent window bf16 and fp32 share, which the kernel guarantees per CTA:
 * B_i <= 110 -> R_i from the bound alone; otherwise R_i is anchored on the exact maximum of the row's first KV tile (valid for
 * B_i - m0 <= 220); a CTA where neither holds runs the running-max path of fgb_attn_fwd_ex. o_peers may be NULL (then `o` is used).
 * fgb_attn_set_stats registers an optional caller-owned device int32[3]; every CTA of a bounded launch adds 1 to
 * counts[mode] (0 = bound only, 1 = first-tile anchored, 2 = running-max fallback). NULL switches the counting off. */
/* fgb_rmsnorm_rope on the q AND the k slice of fused q|k|v rows (columns [0, dim) and [dim, 2*dim) of qkv, row stride ld) plus
 * fgb_head_norm_max of the finished k, in one pass: the single-GPU self-attention prologue (DIT:140-144). */
int fgb_qk_norm_rope(fgb_ctx* ctx, void* qkv, int64_t ld, int32_t rows, int32_t dim, float eps, const void* wq, const void* wk,
                     const void* rope_tab, int32_t gf, int32_t gh, int32_t gw, int32_t token_offset, void* kmax2_f32, void* qmax2_f32 /* optional: the same maximum for q */, void* stream);
int fgb_head_norm_max(fgb_ctx* ctx, const void* x, int64_t ldx, int32_t rows, int32_t heads, void* out_f32, void* stream);
int fgb_attn_set_stats(fgb_ctx* ctx, void* counts_dev_i32x3);
int fgb_attn_fwd_bounded(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale, const void* kmax2,
                         void* lse, int64_t ld_lse, void* workspace, int64_t workspace_bytes, void* const* o_peers,
                         int32_t n_peers, int32_t rows_per_peer, int32_t col_offset, void* stream);
/* The same with qmax2[head] = max_i ||q[i, head]||^2 (optional, NULL = fgb_attn_fwd_bounded): a head whose bound
 * sqrt(qmax2·kmax2)·scale·log2e <= 110 uses ONE reference for all its rows, so a work item starts without the per-row norm pass
 * over the Q tile and without the CTA vote (1.3 us of a 9.5 us cross-attention item). fgb_qk_norm_rope / fgb_rmsnorm_hmax /
 * fgb_recv_norm_rope leave qmax2 as a by-product; fgb_head_norm_max on q gives it to a standalone caller. */
int fgb_attn_fwd_bounded_qk(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale, const void* kmax2,
                            const void* qmax2, void* lse, int64_t ld_lse, void* workspace, int64_t workspace_bytes,
                            void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, int32_t col_offset, void* stream);
/* fgb_rmsnorm_rope without RoPE (the cross-attention query, DIT:176, 99-110) that also leaves hmax2[h] = max over rows of the
 * squared norm of the OUTPUT head slice (fp32 [dim/128], zeroed by the call). */
int fgb_rmsnorm_hmax(fgb_ctx* ctx, void* x, int64_t ldx, int32_t rows, int32_t dim, float eps, const void* weight, void* hmax2_f32,
                     void* stream);

/* Backward of fgb_attn_fwd(_ex): given dout = dL/do, writes dq [s_q, heads*128], dk and dv [s_kv, heads*128] (bf16).
 * The reference obtains this from torch autograd through flash_attention (DIT:27-60) during the stage-2 LoRA
 * fine-tune (config 5; diffusion/loss.py:17-20 -> model_fn -> DiTBlock, PIPE:1348-1360 re-computes each block).
 * o and lse are the forward's outputs; lse and the scratch `delta` are fp32 [heads][ld_stat] with ld_stat a multiple
 * of 64 and >= s_q (delta = rowsum(dout*o) is computed here). Deterministic: two tensor-core kernels (dK|dV with the
 * keys resident, dQ with the queries resident), no atomics. */
int fgb_attn_bwd(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                 const void* o, int64_t ldo, const void* dout, int64_t ld_do, const void* lse, void* delta,
                 int64_t ld_stat, void* dq, int64_t ld_dq, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv,
                 int32_t s_q, int32_t s_kv, int32_t heads, float scale, void* stream);

/* ---- fused memory-bound kernels ------------------------------------------------------------ */

/* y = LayerNorm(x; eps, no affine) * (1 + scale[row]) + shift[row]; rows < rows_mod0 use
 * (shift0, scale0), the rest (shift1, scale1); each [dim] bf16.
 * Replaces norm1/norm2 + modulate (DIT:205-206, 63-64, 224, 227) and Head's norm+modulate
 * (DIT:257, 262-267) without materialising the (1,S,6,D) t_mod of PIPE:1228 / DIT:217-223. */
int fgb_ln_modulate(fgb_ctx* ctx, const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t dim,
                    float eps, const void* shift0, const void* scale0, const void* shift1, const void* scale1,
                    int32_t rows_mod0, void* stream);

/* y = LayerNorm(x; eps) * weight + bias   (norm3, DIT:207, 226). */
int fgb_ln_affine(fgb_ctx* ctx, const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t dim,
                  float eps, const void* weight, const void* bias, void* stream);

/* In place: x = RMSNorm(x over the full row of `dim`; eps) * weight, then (if rope_tab != NULL) the
 * 3-D RoPE of rope_apply on adjacent pairs within each 128-wide head.
 * rope_tab: float2 [1024][64] (cos, sin) — lanes [0,22) rotate with the frame index, [22,43) with the
 * row index, [43,64) with the column index (precompute_freqs_cis_3d, DIT:74-88, 325); token t of
 * this call is global token (token_offset + t) in (f h w) order; tokens >= gf*gh*gw are left
 * un-rotated (the ones-padding of xdit_context_parallel.py:30-55).
 * Replaces RMSNorm (DIT:99-110, 140-141, 176-177) + rope_apply (DIT:91-96, 143-144). */
int fgb_rmsnorm_rope(fgb_ctx* ctx, void* x, int64_t ldx, int32_t rows, int32_t dim, float eps, const void* weight,
                     const void* rope_tab, int32_t gf, int32_t gh, int32_t gw, int32_t token_offset, void* stream);

/* rows[t, c*4 + y*2 + z] = latents[c, f, 2h+y, 2w+z] for token t = (f h w): the im2row side of the
 * Conv3d(k=s=(1,2,2)) patch embedding (DIT:305, 338-344; PIPE:1253, 1260-1261). latents is
 * [channels, gf, 2*gh, 2*gw] contiguous; writes tokens [token_offset, token_offset+rows). */
int fgb_patchify_rows(fgb_ctx* ctx, const void* latents, void* rows_out, int64_t ld_rows, int32_t channels,
                      int32_t gf, int32_t gh, int32_t gw, int32_t token_offset, int32_t rows, void* stream);

/* out[c, f, 2h+y, 2w+z] = head_rows[t, y*2*channels + z*channels + c]  (unpatchify, DIT:346-351). */
int fgb_unpatchify(fgb_ctx* ctx, const void* head_rows, int64_t ld_rows, void* out, int32_t channels, int32_t gf,
                   int32_t gh, int32_t gw, void* stream);

/* One fused denoising update on the latent [channels, frames, hw] (bf16, in place):
 *   n   = noise_neg + cfg_scale * (noise_pos - noise_neg)     (PIPE:302; noise_neg NULL -> n = noise_pos)
 *   x   = x + n * sigma_delta                                  (FlowMatchScheduler.step, FM:144-154)
 *   x[:, 0] = first_frame                                      (PIPE:308-309; first_frame NULL -> skip)
 * with the reference's bf16 rounding after every tensor op. */
int fgb_cfg_fm_step(fgb_ctx* ctx, void* latents, const void* noise_pos, const void* noise_neg,
                    const void* first_frame, float cfg_scale, float sigma_delta, int32_t channels, int32_t frames,
                    int32_t hw, void* stream);

/* out[r, :] = [cos(t_r * w_i), sin(t_r * w_i)], w_i = 10000^(-i/(dim/2)), computed in fp64 and
 * rounded to bf16 (sinusoidal_embedding_1d, DIT:67-71). timesteps: fp32 [rows]. */
int fgb_sinusoidal_embedding(fgb_ctx* ctx, const void* timesteps_f32, void* out, int32_t rows, int32_t dim,
                             void* stream);

/* y = silu(x) elementwise, bf16 (nn.SiLU in time_embedding/time_projection, DIT:312-318). */
int fgb_silu(fgb_ctx* ctx, const void* x, void* y, int64_t n, void* stream);

/* out[r, j] = bf16(a[r, j] + b[j % period])  — `modulation + t_mod` (DIT:217-218, 263, 266). */
int fgb_add_bcast(fgb_ctx* ctx, const void* a, const void* b, void* out, int64_t rows, int64_t cols,
                  int64_t period, void* stream);

/* ---- Ulysses sequence-parallel re-partition (xdit_context_parallel.py:125-146) --------------
 * Pack/unpack between the token-sharded layout x[s_local][groups][heads][128] (groups = 3 for a fused
 * q|k|v row, 1 for the attention output) and the peer-major exchange buffer
 * [world][s_local][groups][heads/world][128] that an all-to-all (NCCL or NVLink peer stores) moves.
 * After the exchange the receive buffer reads as [world*s_local tokens][groups][heads/world][128]:
 * the full (padded) sequence for this rank's heads, directly consumable by fgb_attn_fwd. */
int fgb_sp_pack_heads(fgb_ctx* ctx, const void* x, int64_t ldx, void* send, int32_t s_local, int32_t heads,
                      int32_t groups, int32_t world, void* stream);
int fgb_sp_unpack_heads(fgb_ctx* ctx, const void* recv, void* x, int64_t ldx, int32_t s_local, int32_t heads,
                        int32_t groups, int32_t world, void* stream);

/* ---- Ulysses exchange over NVLink peer memory (no NCCL on the data path) --------------------------------------------
 * One process per GPU; each rank exports the CUDA IPC handle of a caller-owned arena and opens its peers' (the 64-byte
 * handles travel over the torch.distributed control plane once). Afterwards the exchange of
 * xdit_context_parallel.py:125-146 is done by the producing kernels themselves:
 *   forward:  fgb_sp_scatter_heads stores every head slice of this rank's q|k|v rows into the owning peer's receive
 *             matrix (the layout fgb_sp_pack_heads + all-to-all would produce);
 *   return:   fgb_attn_fwd_scatter — the attention epilogue stores each output row straight into the token-major `o`
 *             buffer of the rank that owns the token (compute and collective in ONE kernel);
 *   fgb_sp_barrier orders the two (system-scope release/acquire flags in the arenas; monotonically increasing epoch). */
int fgb_ipc_export(fgb_ctx* ctx, const void* dev_ptr, void* handle_out_64B, int64_t* offset_out);
int fgb_ipc_open(fgb_ctx* ctx, const void* handle_64B, int64_t offset, void** peer_ptr);
int fgb_ipc_close(fgb_ctx* ctx, void* peer_ptr, int64_t offset);
/* peer_bufs / peer_flags / o_peers: HOST arrays of `world` device pointers (index = rank in the SP group, own rank included). */
/* x holds `groups` consecutive groups (group_first .. group_first+groups-1 of the receive matrix's groups_total). */
int fgb_sp_scatter_heads(fgb_ctx* ctx, const void* x, int64_t ldx, void* const* peer_bufs, int32_t s_local, int32_t heads,
                         int32_t groups, int32_t group_first, int32_t groups_total, int32_t world, int32_t rank, void* stream);
/* fgb_rmsnorm_rope whose normalised / rotated rows are not written back but stored head by head into group `group` of
 * the owning peers' receive matrices: q and k of DIT:140-144 leave for the Ulysses exchange straight from registers. */
int fgb_rmsnorm_rope_scatter(fgb_ctx* ctx, const void* x, int64_t ldx, int32_t rows, int32_t dim, float eps, const void* weight,
                             const void* rope_tab, int32_t gf, int32_t gh, int32_t gw, int32_t token_offset,
                             void* const* peer_bufs, int32_t world, int32_t rank, int32_t group, int32_t groups_total,
                             void* stream);
/* Backward of the exchange: x [s_pad, groups*(heads/world)*128] (this rank's heads, all tokens, e.g. dq|dk|dv) -> rows
 * [rank*..] of every token owner's [rows, groups*heads*128] matrix (peer_bufs, row stride ld_dst). */
int fgb_sp_return_heads(fgb_ctx* ctx, const void* x, int64_t ldx, void* const* peer_bufs, int64_t ld_dst, int32_t rows,
                        int32_t s_pad, int32_t heads, int32_t groups, int32_t world, int32_t rank, void* stream);
int fgb_sp_barrier(fgb_ctx* ctx, void* const* peer_flags, int32_t world, int32_t rank, int32_t epoch, void* stream);
/* The same barrier reporting a dead peer instead of trapping: a peer that has not published `epoch` within timeout_clocks SM
 * clocks (0 = the default, ~30 s) makes the waiting thread write `epoch` to *status (device int32, zero while healthy) and
 * return (later barriers of the same exchange see the word set and stop waiting after ~1000 polls; the first epoch stays in
 * it); the kernels behind it then run on stale peer data, so the host must look at *status before it trusts a result
 * (fairygen_b200.sp.SequenceParallel.check). fgb_sp_barrier = this call with status NULL: no way to report, the kernel traps
 * (the stream then fails with a launch error at the next call) rather than hang the box. */
int fgb_sp_barrier_status(fgb_ctx* ctx, void* const* peer_flags, int32_t world, int32_t rank, int32_t epoch, void* status,
                          int64_t timeout_clocks, void* stream);

/* The same exchange with the SEND side fused into the q|k|v projection (DIT:140-142 + USP:125-146) — the default on GPUs:
 *   fgb_gemm_qkv_scatter  out = a[m,k] · w[3*dim,k]ᵀ + bias, never written locally: the epilogue of the 2-CTA tcgen05 GEMM TMA-stores
 *                         every 32 x 64 block straight into the receive matrix of the peer that owns the head (peer_recv[q] =
 *                         base of peer q's [world*m, 3*(heads/world)*128] matrix; this rank writes rows [rank*m, rank*m + m)), so
 *                         the NVLink traffic overlaps the tensor-core main loop. rowsq (fp32 [2][m], zero on entry) receives the
 *                         sum of squares of every q and k row: RMSNorm (DIT:99-110) is over the FULL row, which no rank holds
 *                         after the head split.
 *   fgb_sp_stats_barrier  barrier 0 of the block with the statistics riding along: pushes rowsq into every peer's stats matrix
 *                         (peer_stats[q] = fp32 [2][s_pad], rows [rank*rows, +rows)), zeroes rowsq and kmax2, then exchanges the
 *                         epoch flags like fgb_sp_barrier_status (default time-out; status NULL = trap on time-out).
 *   fgb_recv_norm_rope    receiver side: RMSNorm with the received full-row statistics, weight slice (wq / wk point at this
 *                         rank's heads), 3-D RoPE (DIT:91-96) on the q and k groups of recv [s_pad, 3*hpr*128], in place; leaves
 *                         kmax2[h] = max over the first `tokens` rows of ||k[t,h]||^2 (the bound fgb_attn_fwd_bounded wants). */
int fgb_gemm_qkv_scatter(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, int32_t m, int32_t dim,
                         int32_t k, void* const* peer_recv, int32_t world, int32_t rank, float* rowsq, void* workspace,
                         int64_t workspace_bytes, void* stream);   /* workspace: as fgb_gemm_bf16_sk, may be NULL */
int fgb_sp_stats_barrier(fgb_ctx* ctx, void* const* peer_flags, void* const* peer_stats, void* rowsq, int32_t rows, int32_t s_pad,
                         void* kmax2, int32_t hpr, int32_t world, int32_t rank, int32_t epoch, void* status, void* stream);
int fgb_recv_norm_rope(fgb_ctx* ctx, void* recv, int32_t s_pad, int32_t tokens, int32_t hpr, const void* stats, int32_t dim, float eps,
                       const void* wq, const void* wk, const void* rope_tab, int32_t gf, int32_t gh, int32_t gw, void* kmax2,
                       void* qmax2 /* optional, zeroed by the call */, void* stream);
/* fgb_attn_fwd_ex whose output row of global token t goes to o_peers[t / rows_per_peer][(t % rows_per_peer) * ldo +
 * col_offset + head*128 ...] (col_offset = rank * heads * 128 for Ulysses). */
int fgb_attn_fwd_scatter(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         void* const* o_peers, int32_t n_peers, int64_t ldo, int32_t rows_per_peer, int32_t col_offset,
                         int32_t s_q, int32_t s_kv, int32_t heads, float scale, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* ---- training (BASELINE config 5: stage-2 motion-LoRA fine-tune step) -------------------------------------------
 * The reference trains the `lora_B2` matrices of the 300 adapted Linears (diffusion/training_module.py:266-352, TMOD)
 * under the flow-matching SFT loss (diffusion/loss.py:5-21, LOSS); every gradient below is what torch autograd
 * computes there. Activation gradients are bf16 (as autograd's), reductions / B2 gradients fp32. */

/* dx[m, n_in] = dy[m, k_out] · W[k_out, n_in] — input gradient of nn.Linear with W as stored ([out, in]); same tcgen05
 * kernel as fgb_gemm_bf16 with W fed as an MN-major operand (no transposed weight copy). W is frozen, so there is no
 * full-rank wgrad on this path (TMOD:279-307). */
int fgb_gemm_dgrad(fgb_ctx* ctx, const void* dy, int64_t ld_dy, const void* w, int64_t ldw, void* dx, int64_t ld_dx,
                   int32_t m, int32_t n_in, int32_t k_out, void* stream);

/* Same plus a rank-r term as extra K-blocks: dx = dy·W + u·A1  (u [m, k2] = dy·B2eff, a1 [k2, n_in] row-major). */
int fgb_gemm_dgrad_ex(fgb_ctx* ctx, const void* dy, int64_t ld_dy, const void* w, int64_t ldw, void* dx, int64_t ld_dx,
                      int32_t m, int32_t n_in, int32_t k_out, const void* u, int64_t ld_u, const void* a1, int64_t ld_a1,
                      int32_t k2, void* stream);

/* Backward of fgb_ln_modulate (affine = 0: g = 1 + scale row, rows < rows_mod0 use g0 else g1) or fgb_ln_affine
 * (affine = 1: g0 = weight): out = dres + LN'(x)ᵀ(dy * g). dres may be NULL, and may alias out. */
int fgb_ln_bwd(fgb_ctx* ctx, const void* x, int64_t ldx, const void* dy, int64_t ld_dy, const void* dres, int64_t ld_dres,
               void* out, int64_t ld_out, int32_t rows, int32_t dim, float eps, const void* g0, const void* g1,
               int32_t rows_mod0, int32_t affine, void* stream);

/* Backward of fgb_rmsnorm_rope, in place on dy: x is the PRE-norm input saved by the forward. */
int fgb_rmsnorm_rope_bwd(fgb_ctx* ctx, const void* x, int64_t ldx, void* dy, int64_t ld_dy, int32_t rows, int32_t dim,
                         float eps, const void* weight, const void* rope_tab, int32_t gf, int32_t gh, int32_t gw,
                         int32_t token_offset, void* stream);

/* h = gelu_tanh(z) and dz = dh * gelu_tanh'(z) (the training forward keeps z, so FFN-1 uses FGB_EPI_BIAS). DIT:208 */
int fgb_gelu_tanh(fgb_ctx* ctx, const void* z, void* h, int64_t n, void* stream);
int fgb_gelu_tanh_bwd(fgb_ctx* ctx, const void* z, const void* dh, void* dz, int64_t n, void* stream);

/* out[r, :] = dx[r, :] * gate[r < rows_gate0 ? 0 : 1][:]  — gradient through GateModule (DIT:192-193). */
int fgb_mul_gate(fgb_ctx* ctx, const void* dx, int64_t ld_dx, void* out, int64_t ld_out, int32_t rows, int32_t dim,
                 const void* gate0, const void* gate1, int32_t rows_gate0, void* stream);

/* W_eff[n,k] = W[n,k] + scaling * sum_r (B1[n,r] + bf16(bf16(B2[n,r]*mask[n,r]) * mask_mul)) * A1[r,k]
 * — the stage-2 forward of TMOD:317-352 folded into one weight per step (b1 or b2 may be NULL; mask uint8 or NULL). */
int fgb_lora_merge(fgb_ctx* ctx, const void* w, int64_t ldw, const void* a1, int64_t lda, const void* b1, const void* b2,
                   const void* mask, float mask_mul, float scaling, void* w_eff, int64_t ld_eff, int32_t n, int32_t k,
                   int32_t rank, void* stream);

/* out[n, r] = scaling * bf16(bf16(B2[n,r]*mask[n,r]) * mask_mul) (TMOD:343-346) into a view with row stride ld_out — the W2
 * operand of fgb_gemm_bf16_ex / the k_out x r matrix of the `u = dy·B2eff` dgrad (block-diagonal for the fused q|k|v GEMM). */
int fgb_lora_b2_eff(fgb_ctx* ctx, const void* b2, const void* mask, float mask_mul, float scaling, void* out, int64_t ld_out,
                    int64_t n_rows, int32_t rank, void* stream);

/* The same for every adapted Linear of the model in ONE launch. table: int64 [n_entries][4] on the device, entry =
 * {element offset into b2_flat / mask_flat, rows, destination pointer, destination row stride (elements)}. */
int fgb_lora_b2_eff_batched(fgb_ctx* ctx, const void* b2_flat, const void* mask_flat, const void* table, int32_t n_entries,
                            int32_t rank, float mask_mul, float scaling, void* stream);

/* db[n, r] += mul * mask[n, r] * sum_s dy[s, n] * t[s, r]   (fp32 accumulate; t = A1·x, [rows, rank] bf16).
 * transpose_out = 1 writes db as [rank, n] instead (mask must be NULL): dA = uᵀ·X of the stage-1 LoRA, called with dy := X,
 * t := u = dY·Beff (training_module.py:200-264). */
int fgb_lora_wgrad(fgb_ctx* ctx, const void* dy, int64_t ld_dy, const void* t, int64_t ld_t, void* db_f32, const void* mask,
                   float mul, int32_t rows, int32_t n, int32_t rank, int32_t transpose_out, void* stream);

/* keep-mask of the weight dropout on B2 (`torch.rand_like(B2) > p`, TMOD:338-346): counter-based, reproducible per seed. */
int fgb_bernoulli_mask(fgb_ctx* ctx, void* out_u8, int64_t n, float drop_prob, uint64_t seed, void* stream);

/* latents = (1 - sigma) * x0 + sigma * noise;  target = noise - x0   (flow_match.py:164-175), bf16 roundings as torch's. */
int fgb_fm_noise_target(fgb_ctx* ctx, const void* x0, const void* noise, float sigma, void* latents, void* target, int64_t n,
                        void* stream);

/* loss = weight * mean((pred - target)^2) in fp32 (LOSS:19-20) and dpred = dloss/dpred (bf16; NULL to skip). */
int fgb_mse_loss_grad(fgb_ctx* ctx, const void* pred, const void* target, float weight, void* loss_f32, void* dpred, int64_t n,
                      void* stream);

/* Adjoint of fgb_unpatchify: d_rows[t, y*2C + z*C + c] = dpred[c, f, 2h+y, 2w+z]. */
int fgb_unpatchify_bwd(fgb_ctx* ctx, const void* dpred, void* d_rows, int64_t ld_rows, int32_t channels, int32_t gf, int32_t gh,
                       int32_t gw, void* stream);

/* AdamW on bf16 parameters with fp32 gradient and moments (the reference's optimizer for lora_B2, train.py). */
int fgb_adamw_step(fgb_ctx* ctx, void* param_bf16, const void* grad_f32, void* m_f32, void* v_f32, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * umT5 text encoder: animation/diffsynth/models/wan_video_text_encoder.py ("TENC"), called at
 * pipelines/wan_video.py:404-412. The Linears (bias-free) run on fgb_gemm_bf16; these are the ops around them.
 * --------------------------------------------------------------------------------------------------------------- */

/* out[r, :] = table[ids[r], :] (ids: int64 on the device).  nn.Embedding token_embedding, TENC:233-234, 246. */
int fgb_embedding_rows(fgb_ctx* ctx, const void* table, int64_t ld_table, int32_t vocab, const void* ids, int32_t n, int32_t dim,
                       void* out, int64_t ldo, void* stream);

/* out = weight * bf16(x * rsqrt(mean(x^2) + eps)).  T5LayerNorm.forward, TENC:33-38. */
int fgb_t5_layer_norm(fgb_ctx* ctx, const void* x, int64_t ldx, void* out, int64_t ldo, int32_t rows, int32_t dim, float eps,
                      const void* weight, void* stream);

/* out[r, j] = gate_fc1[r, F + j] * gelu_tanh(gate_fc1[r, j]): the product of T5FeedForward.forward (TENC:18-22, 109) on the
 * output of ONE GEMM against the row-concatenated [gate.0.weight; fc1.weight]. */
int fgb_geglu(fgb_ctx* ctx, const void* gate_fc1, int64_t ld, void* out, int64_t ldo, int32_t rows, int32_t ffn_dim, void* stream);

/* out[h, i] = emb[bucket_of_rel[i], h] (fp32 [heads, n_rel]): T5RelativeEmbedding.forward (TENC:159-169) stored by
 * relative position i = (key - query) + (s_q - 1) instead of as a dense [heads, s_q, s_kv] tensor; bucket_of_rel (int32 on
 * the device) is the host-side restatement of _relative_position_bucket (TENC:171-193). */
int fgb_t5_bias_table(fgb_ctx* ctx, const void* emb, const void* bucket_of_rel, int32_t n_rel, int32_t heads, int32_t buckets,
                      void* out, void* stream);

/* T5Attention.forward core (TENC:74-88), head_dim 64, no score scaling: for every sample b and head h
 *   o[b*s_q + i, h*64:(h+1)*64] = softmax_j(q_i·k_j + bias[h, j - i + s_q - 1], keys with key_mask[b, j] == 0 excluded) · v.
 * q rows b*s_q.., k / v rows b*s_kv..; bias (fp32, from fgb_t5_bias_table) and key_mask (uint8 [batch, s_kv]) may be NULL.
 * A sample whose keys are all masked yields zeros (the host mirror rejects such masks). */
int fgb_t5_attention(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, int32_t batch, int32_t s_q, int32_t s_kv, int32_t heads, const void* bias, const void* key_mask,
                     void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Wan2.2 VAE38 decoder: animation/diffsynth/models/wan_video_vae.py ("VAE"), called at
 * pipelines/wan_video.py:322-323. Feature maps are channels-last zero-bordered grids G[t][H+2][W+2][Cp] (Cp = channels rounded
 * up to 64, padding channels zero), flattened to rows.
 * --------------------------------------------------------------------------------------------------------------- */

/* Convolution as a GEMM over shifted rows: out[r, :] = epilogue(sum_tap x[a_row0 + r + tap_offsets[tap], :] · w[:, tap*cin:(tap+1)*cin]ᵀ
 * + bias), r in [0, m). x: [x_rows, cin] rows (rows outside [0, x_rows) read as zero), w: [n, taps*cin], cin % 64 == 0.
 * grid_h, grid_w > 0: output rows are positions of [.., grid_h, grid_w] grids and the 1-wide border positions are written as
 * zero (the zero padding of the next convolution). epilogue: FGB_EPI_BIAS or FGB_EPI_RESIDUAL (out += ...).
 * Replaces CausalConv3d / nn.Conv2d / 1x1 convs of the decoder (VAE:33-52, 98-105, 274-281, 315-316): a causal 3x3x3 conv has 27
 * taps with offsets (dt*grid_h + dy)*grid_w + dx, dt in {-2,-1,0}, the cached frames (VAE:288-301) stored in front of x. */
int fgb_conv_taps_bf16(fgb_ctx* ctx, const void* x, int64_t ldx, int64_t x_rows, int64_t a_row0, const void* w, int64_t ldw,
                       const void* bias, void* out, int64_t ldo, int32_t m, int32_t n, int32_t cin, int32_t taps,
                       const int32_t* tap_offsets, int32_t grid_h, int32_t grid_w, int32_t epilogue, void* stream);

/* grid interior <- z[c, t, y, x] / inv_std[c] + mean[c]  (latents bf16 [C, T, H, W]; VideoVAE38_.decode, VAE:1328-1331). */
int fgb_vae_latent_rows(fgb_ctx* ctx, const void* z, const void* mean_f32, const void* inv_std_f32, void* grid, int32_t channels,
                        int32_t frames, int32_t h, int32_t w, int32_t cp, void* stream);

/* out = silu?(x / max(|x|_2, 1e-12) * sqrt(channels) * gamma) per row: RMS_norm (VAE:67-70) + nn.SiLU (VAE:274-277, 886). */
int fgb_vae_norm_silu(fgb_ctx* ctx, const void* x, void* out, int64_t rows, int32_t channels, int32_t cp, const void* gamma,
                      int32_t silu, void* stream);

/* nearest-exact x2 in h, w into the interior of a grid of twice the size (Upsample, VAE:73-79). halves = 2: src rows hold
 * 2*cp channels and frame t' of dst is channel half t' % 2 of src frame t' / 2 (temporal up-sampling, VAE:147-156). */
int fgb_vae_upsample2x(fgb_ctx* ctx, const void* src, void* dst, int32_t cp, int32_t frames_dst, int32_t h, int32_t w, int32_t halves,
                       void* stream);

/* main += DupUp3D(x): the parameter-free shortcut of Up_ResidualBlock (VAE:417-439, 506-514). x grid [T, h, w, cin_p], main grid
 * [frames_out, 2h, 2w, cout_p]; first_chunk drops the first factor_t - 1 frames. */
int fgb_vae_dup_up_add(fgb_ctx* ctx, const void* x, void* main, int32_t cin, int32_t cin_p, int32_t cout, int32_t cout_p,
                       int32_t factor_t, int32_t first_chunk, int32_t frames_out, int32_t h, int32_t w, void* stream);

/* In place on bf16 scores [rows, ld]: softmax(scale * s) over the interior positions of one [grid_h, grid_w] frame (columns that
 * are border positions or >= grid_h*grid_w get probability 0).  AttentionBlock.forward, VAE:331-337. */
int fgb_vae_attn_softmax(fgb_ctx* ctx, void* scores, int64_t ld, int32_t rows, int32_t n_cols, int32_t grid_h, int32_t grid_w, float scale,
                         void* stream);

/* Un-patchify (VAE:214-224) of the head grid [frames, h, w, cp] (12 channels used) into the fp32 video [3, video_frames, video_h,
 * video_w] at (t0, y0, x0). weight_f32 == NULL: values = clamp(v, -1, 1) (single_decode, VAE:1212-1215). Otherwise the blending
 * of tiled_decode (VAE:1081-1100, 1128-1150): values += v * mask, weight += mask; bounds_tblr bit 3/2/1/0 = the tile touches the
 * top / bottom / left / right edge of the video (no ramp there); border_y / border_x = ramp widths in pixels. */
int fgb_vae_unpatchify(fgb_ctx* ctx, const void* head, int32_t frames, int32_t h, int32_t w, int32_t cp, void* values_f32,
                       void* weight_f32, int32_t t0, int32_t y0, int32_t x0, int32_t video_frames, int32_t video_h, int32_t video_w,
                       int32_t bounds_tblr, int32_t border_y, int32_t border_x, void* stream);

/* values = clamp(values / weight, -1, 1) (VAE:1151-1152); weight has one plane, values `channels` planes. */
int fgb_vae_blend_finish(fgb_ctx* ctx, void* values_f32, const void* weight_f32, int64_t plane, int32_t channels, void* stream);

/* --- VAE38 encoder side (VideoVAE38_.encode, VAE:1298-1323; Encoder3d_38, VAE:620-733) --- */

/* patchify 'b c f (h q) (w r) -> b (c r q) f h w' (VAE:199-211): video bf16 [3, frames, h, w] -> grid [frames, h/2, w/2, cp]. */
int fgb_vae_patchify_rows(fgb_ctx* ctx, const void* video, void* grid, int32_t frames, int32_t h, int32_t w, int32_t cp, void* stream);

/* dst[t, y, x, (py*2+px)*cp + c] = src[t, 2y+py, 2x+px, c] (grids [frames, h, w, cp] -> [frames, h/2, w/2, 4cp]): turns the stride-2
 * 3x3 convolution with ZeroPad2d((0,1,0,1)) of Resample 'downsample2d/3d' (VAE:106-117, 240-249) into a 2x2-tap stride-1 tap GEMM. */
int fgb_vae_space_to_depth(fgb_ctx* ctx, const void* src, void* dst, int32_t cp, int32_t frames, int32_t h, int32_t w, void* stream);

/* main += AvgDown3D(x): the parameter-free shortcut of Down_ResidualBlock (VAE:363-395, 469-474). x grid [T_in, h_out*fs, w_out*fs,
 * cin_p], main grid [frames_out, h_out, w_out, cout_p]; pad_front zero frames precede x (frame count not a multiple of factor_t). */
int fgb_vae_avg_down_add(fgb_ctx* ctx, const void* x, void* main, int32_t cin, int32_t cin_p, int32_t cout, int32_t cout_p,
                         int32_t factor_t, int32_t factor_s, int32_t pad_front, int32_t frames_out, int32_t h_out, int32_t w_out,
                         void* stream);

/* Latent mean (first z_dim channels of the conv1 grid), normalised (mu - mean) * inv_std (VAE:1313-1320), into fp32
 * [z_dim, out_frames, out_h, out_w] at (t0, y0, x0); weight_f32 != NULL: blended like tiled_encode (VAE:1181-1203), arguments as in
 * fgb_vae_unpatchify. fgb_vae_blend_divide then forms values / weight (no clamp). */
int fgb_vae_latent_out(fgb_ctx* ctx, const void* grid, int32_t frames, int32_t h, int32_t w, int32_t cp, const void* mean_f32,
                       const void* inv_std_f32, int32_t z_dim, void* values_f32, void* weight_f32, int32_t t0, int32_t y0, int32_t x0,
                       int32_t out_frames, int32_t out_h, int32_t out_w, int32_t bounds_tblr, int32_t border_y, int32_t border_x,
                       void* stream);
int fgb_vae_blend_divide(fgb_ctx* ctx, void* values_f32, const void* weight_f32, int64_t plane, int32_t channels, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FAIRYGEN_B200_H_ */
