/*
 * stgcma.h -- C ABI of libstgcma.so: the B200 (sm_100a) kernels behind the STG-CMA
 * audio-visual adapter TRAIN STEP (BASELINE.json north_star).
 *
 * The reference (kaiw7/STG-CMA) is pure PyTorch and owns no kernels: every entry point
 * below replaces a group of stock torch call sites inside
 *     AVE/model/Swin_AVE.py  SwinTransformer2D_Adapter_New.forward(a, v, 'fusion')  :1565-1599
 * and the autograd graph torch would build for them under the caller's freeze rule
 * (AVE/traintest_adapt_ave29.py:38-61).  The reference-side binding is the ctypes stub in
 * INTEGRATION.md; the shipped host side is stg-cma_b200/stgcma/lib.py.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless stated otherwise.
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated or freed
 *     here; all work is enqueued on `stream` (a cudaStream_t passed as void*), no implicit sync.
 *   - return value: 0 = ok, otherwise an STGCMA_E* code; stgcma_last_error() describes it
 *     (thread local).  No C++ exception crosses the ABI.
 *   - `dtype` selects the ACTIVATION storage type: STGCMA_F32 is the fp32 check mode (all math
 *     fp32 on CUDA cores), STGCMA_BF16 the production mode (bf16 storage, fp32 accumulate,
 *     tensor cores).  Frozen parameters used by the element-wise kernels (LayerNorm affine,
 *     bias tables, biases) are always fp32 (the caller's master weights); GEMM weight operands
 *     use the activation dtype.
 *   - activations are token-major [rows, channels]; "rows" of the two modalities are stacked
 *     video first, audio second, because both streams share the frozen weights.
 *   - refuses to run (STGCMA_EARCH) on anything that is not compute capability 10.0 (the cubins are sm_100a).
 */
#ifndef STGCMA_H_
#define STGCMA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { STGCMA_F32 = 0, STGCMA_BF16 = 1 };

enum {
  STGCMA_OK = 0,
  STGCMA_EINVAL = 1,   /* bad argument (shape / alignment / unsupported size) */
  STGCMA_ECUDA = 2,    /* a CUDA runtime / driver call failed                 */
  STGCMA_EARCH = 3     /* device is not sm_100                                */
};

/* GEMM epilogues: v = rowscale[m] * (acc + bias[n])   (rowscale optional, fp32 [M])       */
enum {
  STGCMA_EPI_LINEAR = 0,  /* out = res1 + res2 + v                 (each term optional)      */
  STGCMA_EPI_GELU = 1,    /* out = gelu_erf(v); out_pre = v when out_pre != NULL            */
  STGCMA_EPI_DGELU = 2,   /* out = v * gelu_erf'(aux[m,n])   (backward through nn.GELU)      */
  STGCMA_EPI_MULAUX = 3,  /* out = v * aux[m,n]              (dropout keep-mask)             */
  STGCMA_EPI_QGELU = 4,   /* out = v * sigmoid(1.702 v) (CLIP's QuickGELU, CLIP_AVE.py:41-43); out_pre = v   */
  STGCMA_EPI_DQGELU = 5,  /* out = v * d/dx[x sigmoid(1.702 x)](aux[m,n])                          */
  /* derivative-saving forms: the forward evaluates the activation AND its derivative (the erf / sigmoid terms are
   * shared) and stores the DERIVATIVE in out_pre; the backward of the layer is then the cheap MULAUX epilogue
   * (out = v * aux) instead of a second transcendental evaluation per element.                                  */
  STGCMA_EPI_GELU_G = 6,  /* out = gelu_erf(v);          out_pre = gelu_erf'(v)                                   */
  STGCMA_EPI_QGELU_G = 7  /* out = v * sigmoid(1.702 v); out_pre = d/dv [v sigmoid(1.702 v)]                      */
};

const char* stgcma_last_error(void);
int stgcma_version(void);
/* 0 when device `dev` can run the library (compute capability 10.0), STGCMA_EARCH otherwise */
int stgcma_check_device(int dev);

/* ---- PatchEmbed3D: Conv3d(k=s=(1,4,4)) + bias + LayerNorm   (Swin_AVE.py:1104-1124) -------
 * img fp32 [B, Cin, T, H, W] (audio: Cin = 1); w fp32 [C, Cin*16] (= proj.weight flattened),
 * bias/gamma/beta fp32 [C]; out [B*T, (H/4)*(W/4), C] in `dtype`, token order (b t)(y x).    */
int stgcma_patch_embed_fwd(const float* img, const float* w, const float* bias, const float* gamma,
                           const float* beta, void* out, int B, int Cin, int T, int H, int W, int C,
                           int dtype, void* stream);

/* ---- nn.LayerNorm(C), eps 1e-5, frozen affine (norm1/norm2/norm; Swin_AVE.py:346,353,1314) --
 * x is the RESIDUAL STREAM and has its own dtype `xdtype` (fp32 in both modes: the reference under the
 * caller's autocast keeps the residual in fp32 too); y is an activation of type `dtype`.
 * merge = 0: x, y are [rows, C].
 * merge = 1: PatchMerging gather fused in (Swin_AVE.py:969-978): x is [frames, R, R, C/4],
 *            y is [frames*(R/2)^2, C] with channel blocks (2y,2x),(2y+1,2x),(2y,2x+1),(2y+1,2x+1);
 *            rows = frames*(R/2)^2.
 * mean / rstd fp32 [rows] are written when non-NULL (saved for backward).                    */
int stgcma_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                         float* rstd, int64_t rows, int C, int merge, int R, int xdtype, int dtype, void* stream);
/* dx = dres + dLN/dx (dy)   -- no dgamma/dbeta: every LayerNorm on this path is frozen.
 * dy has type `dydtype`; x, dres, dx have type `xdtype` (stream); dx_copy (optional) receives the same
 * values in the activation type `dtype` -- the copy the next GEMM consumes as its A operand.
 * merge = 1 scatters back to [frames, R, R, C/4] (dres must then be NULL).                   */
int stgcma_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                         const float* gamma, const void* dres, void* dx, void* dx_copy, int64_t rows, int C,
                         int merge, int R, int dydtype, int xdtype, int dtype, void* stream);

/* ---- nn.Linear family: out[m, n] = epilogue( sum_k A[m,k] * W[g][n,k] + bias[g][n] ) --------
 * Replaces attn.qkv / attn.proj / mlp.fc1 / mlp.fc2 / downsample.reduction / adapter D_fc1,
 * D_fc2 / mlp_head (Swin_AVE.py:203,205,116,118,957,15-16,1320-1322) and, with the transposed
 * weight, their dX backward.  A [M, K] (row stride lda), W [groups, N, K] (row stride ldw,
 * group stride N*ldw), bias fp32 [groups, N] or NULL.  Rows are split in `groups` equal parts of
 * M/groups rows, part g using W[g] (groups = 2: per-modality adapter weights).
 * bf16 mode: tcgen05/TMEM kernel fed by TMA; needs K, lda, ldw multiples of 8 and 16-byte
 * aligned A / W; N multiple of 8.  res1/res2/aux/out/out_pre share row stride ldo.
 * Skinny contractions with a residual (K <= 128, K % 16 == 0, N % 16 == 0, res1 != NULL, LINEAR:
 * the adapters' D_fc2 + residual adds and the dX of D_fc1) are HBM streams and run in a warp-level
 * streaming kernel instead (gemm_skinny.cu); same arguments, same results.
 * out_dtype: type of `out` and `res1` -- `dtype`, or STGCMA_F32 when they are the fp32 residual
 * stream (x_new = x + branch); res2 / aux / out_pre always have the activation type.          */
int stgcma_gemm(const void* A, const void* W, const float* bias, const void* res1, const void* res2,
                const float* rowscale, const void* aux, void* out, void* out_pre, int64_t M, int N, int K,
                int64_t lda, int64_t ldw, int64_t ldo, int groups, int epilogue, int dtype, int out_dtype,
                void* stream);

/* ---- fused frozen MLP of the (G) sub-block, bf16 production mode only (Swin_AVE.py:790-797) ------
 * fwd:  xn = LayerNorm(x; gamma, beta)  (norm2 :353)         x: fp32 residual stream [M, C]
 *       y  = fc2(GELU(fc1(xn)))         (Mlp.forward :121-127) W1 [H, C], b1 [H], W2 [C, H], b2 [C]
 *       hpre = y Wa[g]^T + ba[g], h = GELU(hpre)  (S_Adapter[_Audio].D_fc1 + act :796-797)  Wa [groups, rp, C]
 *       The H-wide hidden activation never leaves the SM (chained tcgen05 GEMMs through TMEM / shared memory).
 *       Writes y [M, C], hpre, h [M, rp] (activation dtype) and mean, rstd fp32 [M] (saved for backward).
 * bwd:  dxn = [ ((dhpre Wa[g] + dxb) W2) * gelu'(xn W1^T + b1) ] W1      -- gradient wrt xn = LN2(x), [M, C]
 *       (the hidden pre-activation is recomputed from x and the saved statistics; nothing H-wide is stored).
 *       dhpre [M, rp]: gradient wrt hpre; dxb [M, C]: gradient wrt y coming from the residual (activation dtype).
 *       Wat [groups, C, rp] = Wa transposed, W2t [H, C] = W2 transposed, W1t [C, H] = W1 transposed.
 * Limits: C % 32 == 0, C <= 384 (fwd) / 256 (bwd); H % 64 == 0; rp % 16 == 0, rp <= 128; all pointers 16-byte
 * aligned.  Returns STGCMA_EINVAL outside them (the caller then composes layernorm + gemm).       */
int stgcma_mlp_fwd(const float* x, const float* gamma, const float* beta, const void* W1, const float* b1,
                   const void* W2, const float* b2, const void* Wa, const float* ba, void* y, void* hpre, void* h,
                   float* mean, float* rstd, int64_t M, int C, int H, int rp, int groups, int dtype, void* stream);
int stgcma_mlp_bwd(const void* dhpre, const void* dxb, const float* x, const float* gamma, const float* beta,
                   const float* mean, const float* rstd, const void* Wat, const void* W1, const float* b1,
                   const void* W2t, const void* W1t, void* dxn, int64_t M, int C, int H, int rp, int groups,
                   int dtype, void* stream);

/* ---- weight gradients of the trainable nn.Linear layers (adapters, head) -------------------
 * dW[i, j] += sum_m rowscale[m] * A[m, i] * B[m, j]   (i < Ka, j < Nb; dW fp32, row stride ldo)
 * csumA[i] += sum_m rowscale[m] * A[m, i]  (when non-NULL);  csumB likewise.  Accumulates
 * (atomics) so the destination is normally a slice of the pre-zeroed flat gradient bucket.
 * bf16 mode (tensor cores): rowscale[m] * A[m, :] is rounded to bf16 before the product, which is
 * exact for the DropPath keep masks {0, 1/keep} up to the rounding of 1/keep (2^-9 relative).   */
int stgcma_gemm_tn(const void* A, const void* B, const float* rowscale, float* dW, float* csumA,
                   float* csumB, int64_t M, int Ka, int Nb, int64_t lda, int64_t ldb, int64_t ldo, int dtype,
                   void* stream);

/* ---- Swin window attention (WindowAttention.forward, temporal=False; Swin_AVE.py:238-276) ---
 * with torch.roll(-shift) + window_partition + window_reverse + roll(+shift) folded into the
 * addressing (Swin_AVE.py:728-740,766-776) and the 0/-100 shift mask (:368-387) computed from
 * coordinates.  qkv [frames, R*R, 3C] channel order [q|k|v] x heads x 32; rpb fp32
 * [(2ws-1)^2, heads]; out [frames, R*R, C].  head_dim must be 32 (true for every Swin).      */
int stgcma_window_attn_fwd(const void* qkv, const float* rpb, void* out, int64_t frames, int R, int ws,
                           int shift, int heads, int dtype, void* stream);
int stgcma_window_attn_bwd(const void* qkv, const float* rpb, const void* dout, void* dqkv, int64_t frames,
                           int R, int ws, int shift, int heads, int dtype, void* stream);

/* ---- temporal attention (WindowAttention.forward, temporal=True; Swin_AVE.py:244-255) -------
 * sequence = the T frames of one (modality, clip, token); qkv [G*B*T, L, 3C]; tbl fp32
 * [G, 2T-1, heads] (G = 2: video table, audio table); out [G*B*T, L, C].  T <= 16.
 * bwd accumulates dtbl fp32 [G, 2T-1, heads] (the tables are trainable).                     */
int stgcma_temporal_attn_fwd(const void* qkv, const float* tbl, void* out, int G, int B, int T, int L,
                             int heads, int dtype, void* stream);
int stgcma_temporal_attn_bwd(const void* qkv, const float* tbl, const void* dout, void* dqkv, float* dtbl,
                             int G, int B, int T, int L, int heads, int dtype, void* stream);

/* ---- cross-modal token attention of the adapters (Swin_AVE.py:753-760 and :801-808) ---------
 * h [2, frames, L, rp]: GELU(D_fc1(.)) features, video block then audio block, rp = hidden
 * width padded to a multiple of 16 (pad columns zero).  For each frame and token group
 *     hv' = hv + gate[0] * softmax(hv ha^T) ha        attn_out[0] = softmax(hv ha^T) ha
 *     ha' = ha + gate[1] * softmax(ha hv^T) hv        attn_out[1] = softmax(ha hv^T) hv
 * (logits unscaled, both directions use the un-updated features).
 *   window variant: groups = the (shifted) ws x ws windows of the R x R token map, L = R*R.
 *   global variant: one group = all L tokens of the frame (flash-style, never materialises
 *                   the L x L scores); also writes lse fp32 [2, frames, L].
 * gate: fp32 [2] on the device.  out, attn_out: [2, frames, L, rp].
 * bwd: dout = grad wrt h'.  Writes dhpre = (grad wrt h) * gelu'(hpre) when hpre != NULL, else
 * grad wrt h; accumulates dgate fp32 [2].  The global variant needs `work`: fp32 [2, frames, L]
 * scratch owned by the caller.                                                                */
int stgcma_xmodal_window_fwd(const void* h, const float* gate, void* out, void* attn_out, int64_t frames,
                             int R, int ws, int shift, int rp, int dtype, void* stream);
int stgcma_xmodal_window_bwd(const void* h, const void* hpre, const float* gate, const void* attn_out,
                             const void* dout, void* dh, float* dgate, int64_t frames, int R, int ws,
                             int shift, int rp, int dtype, void* stream);
/* global variant with different row counts per modality (CLIP: 197 video / 49 audio tokens per frame,
 * CLIP_AVE.py:380-397): modality 0 has L rows per frame, modality 1 has La (La <= 0 means La = L); the
 * tensors are then the two blocks stacked: [frames, L, rp] followed by [frames, La, rp]; lse / work
 * likewise [frames * L] followed by [frames * La].                                                  */
int stgcma_xmodal_global_fwd(const void* h, const float* gate, void* out, void* attn_out, float* lse,
                             int64_t frames, int L, int La, int rp, int dtype, void* stream);
int stgcma_xmodal_global_bwd(const void* h, const void* hpre, const float* gate, const void* attn_out,
                             const float* lse, const void* dout, void* dh, float* dgate, float* work,
                             int64_t frames, int L, int La, int rp, int dtype, void* stream);

/* ---- head pooling: mean over the N tokens of a frame, cat(audio, video) (Swin_AVE.py:1588-1596)
 * xn [2, frames, N, C] (video, audio) -> feat [frames, 2C] = [audio mean | video mean].      */
int stgcma_head_pool_fwd(const void* xn, void* feat, int64_t frames, int N, int C, int dtype, void* stream);
int stgcma_head_pool_bwd(const void* dfeat, void* dxn, int64_t frames, int N, int C, int dtype, void* stream);

/* ---- per-step refresh of the compute copies of the TRAINABLE parameters ---------------------
 * The optimizer updates the caller's fp32 master nn.Parameters in place; the kernels consume
 * padded / transposed copies in the activation dtype.  One launch executes a device-resident
 * table of 2-D copies: dst[c*dst_ld + r] (transpose) or dst[r*dst_ld + c] = src[r*src_ld + c].
 * Padding regions of dst are never written (they stay zero).                                 */
typedef struct StgcmaPackEntry {
  const float* src;
  void* dst;
  int32_t rows, cols;       /* of src */
  int64_t src_ld, dst_ld;   /* in elements */
  int32_t transpose;
  int32_t dst_f32;          /* 1: dst is fp32 regardless of dtype (biases, tables) */
} StgcmaPackEntry;
int stgcma_pack_params(const StgcmaPackEntry* table_dev, int n, int dtype, void* stream);

/* ---- optimizer tail: Adam over every trainable tensor in ONE launch -----------
 * Replaces torch.optim.Adam.step() of the caller (AVE/traintest_adapt_ave29.py:68,165-168: two
 * parameter groups, betas (0.95, 0.999), weight_decay 5e-7 applied as L2 on the gradient, no
 * amsgrad) -- ~455 tensors, i.e. hundreds of small multi-tensor launches per step in stock torch.
 * Per element, with g = grad_scale * grad + wd * p:
 *     m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;
 *     p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)          bc = 1 - beta^step (host computed)
 * `table_dev` is a device array of n entries; hyper-parameters are per GROUP (<= 8), passed by value.
 * All tensors fp32 (the caller's master weights and the flat gradient bucket).                    */
#define STGCMA_ADAM_MAX_GROUPS 8
typedef struct StgcmaAdamEntry {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  int32_t group;
  int32_t pad_;
} StgcmaAdamEntry;
typedef struct StgcmaAdamGroups {
  float lr[STGCMA_ADAM_MAX_GROUPS], weight_decay[STGCMA_ADAM_MAX_GROUPS], beta1[STGCMA_ADAM_MAX_GROUPS],
      beta2[STGCMA_ADAM_MAX_GROUPS], eps[STGCMA_ADAM_MAX_GROUPS], bias_corr1[STGCMA_ADAM_MAX_GROUPS],
      bias_corr2_sqrt[STGCMA_ADAM_MAX_GROUPS],
      one_minus_beta1[STGCMA_ADAM_MAX_GROUPS], one_minus_beta2[STGCMA_ADAM_MAX_GROUPS];  /* rounded from double, as torch does */
} StgcmaAdamGroups;
int stgcma_adam_step(const StgcmaAdamEntry* table_dev, int n, const StgcmaAdamGroups* groups_host, float grad_scale,
                     void* stream);
/* The same step under torch.amp.GradScaler (the caller's loop, traintest_adapt_ave29.py:158-168) WITHOUT the scaler's
 * separate unscale pass and WITHOUT its per-step host read of found_inf: `scale_dev` (the scaler's current scale) and
 * `found_inf_dev` (1.0 when a gradient is inf / nan) are DEVICE scalars read by the kernel, as in torch's own fused Adam
 * (the optimizer advertises `_step_supports_amp_scaling`).  Gradients are divided by *scale_dev and written back
 * un-scaled; when *found_inf_dev != 0 nothing is updated and *skipped_dev (int32, cumulative) is incremented, so the
 * bias corrections -- computed in the kernel from steps_host[group] - *skipped_dev -- do not advance either.
 * groups_host->bias_corr1 / bias_corr2_sqrt are ignored.                                                           */
/* ---- data-parallel exchange: mean of the flat gradient bucket over the GPUs of one box --------------
 * Replaces the caller's nn.DataParallel reduce / an NCCL all-reduce of the adapter + head gradients
 * (traintest_adapt_ave29.py:32-33): every rank's bucket lives in symmetric (peer-mapped) memory; each rank reads the W
 * buckets over NVLink / NVSwitch with 128-bit loads and writes  out[i] = scale * sum_r peer[r][i]  into its own local
 * tensor -- a one-shot all-reduce (8.9 MB for Swin-T: latency bound, one launch, no staging copies).  `peers_dev` is a
 * DEVICE array of `world` pointers to the ranks' buckets (this rank's own included), all 16-byte aligned; n floats.
 * `mc` (nullable): the NVLS multicast address of the same symmetric allocation; when given, the sum is taken by the
 * switch (multimem.ld_reduce) and `peers_dev` is not read.  The caller brackets the launch with cross-rank barriers
 * (the buckets must be complete before, and must not be rewritten until every rank has read them).                   */
int stgcma_bucket_allreduce(const float* const* peers_dev, int world, const float* mc, float* out, int64_t n, float scale,
                            void* stream);

/* Two-shot variant for larger boxes (the one-shot reads W buckets per rank): rank r reduces ONLY its slice
 * [r n / W, (r + 1) n / W) of the bucket through the switch (multimem.ld_reduce on `mc_in`) and broadcasts the scaled result
 * into every rank's copy of a second symmetric buffer with multimem.st on `mc_out` -- n / W elements per rank instead of n.
 * Barriers by the caller before (buckets complete) and after (every slice has landed everywhere).  n % (4 world) == 0.  */
int stgcma_bucket_allreduce_nvls(const float* mc_in, float* mc_out, int64_t n, int rank, int world, float scale, void* stream);

int stgcma_adam_step_amp(const StgcmaAdamEntry* table_dev, int n, const StgcmaAdamGroups* groups_host,
                         const int32_t* steps_host /* [STGCMA_ADAM_MAX_GROUPS] */, const float* scale_dev,
                         const float* found_inf_dev, int32_t* skipped_dev, void* stream);

/* ==== CLIP-ViT variant of the step (MM_CLIP_AVE, AVE/model/CLIP_AVE.py:716; BASELINE.json configs[1]) ================

 * ---- multi-head attention core of nn.MultiheadAttention (CLIP_AVE.py:54, :106-108; no mask, no dropout) -------------
 * o[:, h] = softmax(q_h k_h^T / sqrt(hd)) v_h per sequence and head; the packed in-projection and the out-projection
 * around it are stgcma_gemm calls.  qkv [rows, 3*heads*hd], channel order [q | k | v] x heads x hd (in_proj_weight
 * layout); out [rows, heads*hd]; lse fp32 [rows, heads] (natural log, saved for backward).
 * Sequences: n_seq sequences of seq_len tokens; token i of sequence s is row
 *     (s / inner) * seq_len * inner + (s % inner) + i * inner
 *   inner = 1: contiguous sequences = the spatial attention over the N tokens of a frame (:385, :388)
 *   inner = N: the temporal attention 'n (b t) d -> t (b n) d' (:369-377): sequence = the T frames of one (clip, token)
 *              of a [clips, T, N] row layout, n_seq = clips * N, seq_len = T.
 * bf16: hd in {16, 32, 64, 96, 128}, sequences must fit shared memory (seq_len <= ~300); fp32: hd <= 128.
 * bwd writes dqkv [rows, 3*heads*hd]; `work` = fp32 [rows, heads] scratch owned by the caller (bf16 mode).          */
int stgcma_mha_fwd(const void* qkv, void* out, float* lse, int64_t n_seq, int seq_len, int inner, int heads, int hd,
                   int dtype, void* stream);
int stgcma_mha_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, float* work,
                   int64_t n_seq, int seq_len, int inner, int heads, int hd, int dtype, void* stream);

/* ---- patchify: conv1 / conv1_audio, kernel = stride = P, no bias (CLIP_AVE.py:728,:741,:1091,:1107) as a gather in
 * front of stgcma_gemm.  img fp32 [B, Cin, T, H, W] (audio: Cin = 1, [B, 1, T, 102, 128]);
 * out[((b*T + t)*Hp + py)*Wp + px][(c*P + ky)*P + kx] = img[b][c][t][py*P + ky][px*P + kx], Hp = (H-P)/P + 1 (rows
 * that do not fill a patch are dropped, like the convolution); row stride ldo >= Cin*P*P elements (pad columns are
 * not written).                                                                                                    */
int stgcma_im2col(const float* img, void* out, int B, int Cin, int T, int H, int W, int P, int64_t ldo, int dtype,
                  void* stream);
/* ---- token assembly (CLIP_AVE.py:1094-1103): x0[bt][0] = class_embedding, x0[bt][1 + p] = patches[bt][p], then
 * + positional_embedding[i] + temporal_embedding[bt % T].  patches [BT*(N-1), D] (`dtype`), cls fp32 [D], pos fp32
 * [N, D], temb fp32 [T, D]; x0 fp32 [BT*N, D] (the residual stream; ln_pre follows as stgcma_layernorm_fwd).        */
int stgcma_clip_embed_fwd(const void* patches, const float* cls, const float* pos, const float* temb, float* x0,
                          int64_t BT, int T, int N, int D, int dtype, void* stream);
/* gradient of the trainable temporal embedding (CLIP_AVE.py:749-752): dtemb[t] += sum_{b, i} dx0[(b*T + t)*N + i]  */
int stgcma_clip_temb_bwd(const float* dx0, float* dtemb, int64_t B, int T, int N, int D, void* stream);
/* ---- head input (CLIP_AVE.py:1124-1137): ln_post (TRAINABLE affine) on the class-token row of every frame, then
 * cat(audio, video).  x fp32 stream: [frames, Nv, D] video block followed by [frames, Na, D] audio block;
 * feat fp32 [frames, 2D] = [LN(x_a[f, 0]) | LN(x_v[f, 0])]; mean / rstd fp32 [2*frames] (video rows then audio rows).
 * bwd: writes ONLY the class-token rows of dx and of its optional copy dx_copy in the activation type `dtype` (the
 * caller zero-fills both), accumulates dgamma / dbeta fp32 [D].                                                    */
int stgcma_cls_ln_fwd(const float* x, const float* gamma, const float* beta, float* feat, float* mean, float* rstd,
                      int64_t frames, int Nv, int Na, int D, void* stream);
int stgcma_cls_ln_bwd(const float* dfeat, const float* x, const float* mean, const float* rstd, const float* gamma,
                      float* dx, void* dx_copy, float* dgamma, float* dbeta, int64_t frames, int Nv, int Na, int D,
                      int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STGCMA_H_ */
