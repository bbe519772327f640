/* rdst_b200.h -- C ABI of librdst_b200.so: hand-written sm_100a kernels for the RDST super-resolution
 * network hot path (GinZhu/RDST: networks/rdst_variations.py + networks/swin_transformer_sr.py).
 *
 * Conventions:
 *   - plain device pointers + explicit shapes/strides, no torch types; the caller owns every buffer;
 *   - every entry point returns 0 on success or a negative RDST_E_* code, never throws;
 *     rdst_last_error() returns a thread-local message for the last failure on the calling thread;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*), never synchronise, allocate
 *     nothing and keep no global state => safe to capture in a CUDA graph; the caller must have the
 *     device that owns the pointers current (the Python wrapper holds a torch.cuda.device guard);
 *   - `dtype` selects the STORAGE type of activations: RDST_F32 or RDST_BF16.  Arithmetic is always
 *     fp32-accumulate; the *_tc entry points use bf16 tcgen05 tensor-core operands.
 *
 * Activation layout ("token-major, padded"): every activation is a 2-D array [T][ld] with T = B*H*W
 * tokens in raster order and channels padded so that GEMM K is a multiple of 16/32:
 *     dense buffer D  : ld = 160 ; trunk x0 at [0,60), growth g_j (30 ch) at [64+32(j-1), +30), pads are 0
 *     STL work buffer : ld = Cp  ; Cp = 64/96/128 for C = 60/90/120 (same channel->position map, truncated)
 *     conv feature map: ld = 64  ; 60 real channels
 * Weights are pre-packed by the host into the same padded K order (zero columns at pads); LayerNorm
 * affine (gamma, beta) and the attention scale are folded into the following linear layer by the host.
 *
 * Each declaration cites the reference code it replaces (paths relative to the reference repo).
 */
#ifndef RDST_B200_H
#define RDST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDST_ABI_VERSION 1

#define RDST_F32  0
#define RDST_BF16 1

#define RDST_OK             0
#define RDST_E_INVALID     -1   /* bad argument (null pointer, unsupported size, misalignment) */
#define RDST_E_CUDA        -2   /* a CUDA runtime call failed; see rdst_last_error()           */
#define RDST_E_UNSUPPORTED -3   /* outside the supported envelope                              */

int         rdst_abi_version(void);
const char* rdst_last_error(void);
/* 1 if the library was built with the tcgen05 (sm_100a) kernels and the current device can run them. */
int         rdst_has_tcgen05(void);

/* ---- generic CUDA-core ("simt") kernels: fp32 mode of the module, and the validation path -------------- */

/* Y[t][n] = out_scale * act( LNhat(X[t][0:K]) . W[n][0:K] + bias[n] ) + R[t][n]
 *   LNhat: if ln_creal > 0, x <- (x - mean) * rstd with statistics over the K stored values divided by
 *          ln_creal real channels (pads are zero), eps 1e-5; gamma/beta pre-folded into W/bias.
 *   act  : 0 none, 1 exact-erf GELU.   R may be NULL.
 * Replaces nn.Linear (+ preceding nn.LayerNorm, + following nn.GELU / residual add):
 *   swin_transformer_sr.py:117 (qkv), :139 (proj), :24-27 (Mlp), :240,:271-272 (norm1/2 + residuals),
 *   rdst_variations.py:309-313,339-340 (DenseSTLayer tail LN+Linear, *dense_scale, in-place "cat"). */
int rdst_linear_fwd(const void* x, int64_t ldx, const float* w, const float* bias,
                    const void* resid, int64_t ldr, void* y, int64_t ldy,
                    int64_t T, int K, int N, int ln_creal, int act, float out_scale,
                    int dtype, void* stream);

/* Windowed multi-head attention core on an 8x8 window grid with optional cyclic shift.
 *   qkv : [T][ldq], q at [0,C), k at [C,2C), v at [2C,3C), head h owns channels [h*C/heads, (h+1)*C/heads);
 *         q is already scaled (scale folded into the qkv weights).
 *   out : [T][ldo], channel order (head, d) as in the reference's transpose(1,2).reshape (:138).
 *   table: relative_position_bias_table (225, heads) fp32; index is (dh+7)*15+(dw+7) (:89-98).
 *   shift: 0 or 4.  The roll(-s)/window_partition/window_reverse/roll(+s) of swin_transformer_sr.py:245-265
 *   is index arithmetic; the {0,-100} mask of calculate_mask (:211-232) is evaluated in closed form.
 * Replaces WindowAttention.forward :120-138 and SwinTransformerBlock.forward :243-268. */
int rdst_window_attention_fwd(const void* qkv, int64_t ldq, const float* table, void* out, int64_t ldo,
                              int B, int H, int W, int C, int heads, int shift, int dtype, void* stream);

/* 3x3 same-padding convolution on token-major (NHWC) data as an implicit GEMM.
 *   x : [B*H*W][ldx] (Cin stored channels), w : [N][9][Cin] (tap = ky*3+kx), bias : [N]
 *   y[t][n] = out_scale * (conv + bias) + R[t][n]                           (shuffle == 0)
 *   shuffle == 2: N = 4*G; output pixel (2y+dy, 2x+dx) of a [B*2H*2W][ldy] map receives channels
 *                 n in [G*(2dy+dx), G*(2dy+dx)+G)  -- nn.PixelShuffle(2) folded into the store.
 * Replaces RDSTB.conv (LFF) + patch_unembed/patch_embed transposes (rdst_variations.py:444-445),
 * conv_after_body (:1349-1350) and UpSampler conv+PixelShuffle (common.py:129-132). */
int rdst_conv3x3_fwd(const void* x, int64_t ldx, const float* w, const float* bias,
                     const void* resid, int64_t ldr, void* y, int64_t ldy,
                     int B, int H, int W, int Cin, int N, float out_scale, int shuffle,
                     int dtype, void* stream);

/* 3x3 same-padding convolution followed by an activation (act: 0 none, 1 GELU, 2 LeakyReLU(0.2)); same layout rules as
 * rdst_conv3x3_fwd, no residual / scale / shuffle:  y[t][n] = act(conv + bias).  Rows n >= the real output channels must be
 * zero in w / bias so that pad channels stay zero.  rdst_linear_fwd accepts the same act codes (1x1 convolution).
 * Replaces the Conv2d + LeakyReLU pairs of RDSTB's '3conv' fusion (rdst_variations.py:422-427). */
int rdst_conv3x3_act_fwd(const void* x, int64_t ldx, const float* w, const float* bias, void* y, int64_t ldy,
                         int B, int H, int W, int Cin, int N, int act, int dtype, void* stream);

/* Shallow feature extraction: img (B,1,H,W) fp32 NCHW -> feat0[t][0:64] = conv3x3(1->60) (kept for the global
 * residual) and dense[t][0:64] = LayerNorm_60(feat0) * gamma + beta (patch_embed.norm), pads zero.
 * `in_scale`/`in_bias` fold sub_mean (MeanShift 1x1).  w : [60][9], bias/gamma/beta : [60].
 * Replaces sub_mean + head + patch_embed (rdst_variations.py:1343-1344,1329; swin_transformer_sr.py:515-519). */
int rdst_head_fwd(const float* img, float in_scale, float in_bias, const float* w, const float* bias,
                  const float* gamma, const float* beta, void* feat0, int64_t ldf, void* dense, int64_t ldd,
                  int B, int H, int W, int dtype, void* stream);

/* Stand-alone LayerNorm over the first `creal` of `ld` stored channels with affine, times out_scale.
 * Output channels >= creal are pads: the bf16 fast path (creal <= 64, 16-byte aligned rows) writes zeros to the pads
 * that share the last 8-channel group with real channels, the generic path leaves all pads untouched.
 * Replaces RDSTSR.norm + *global_res_scale (rdst_variations.py:1337,1347). */
int rdst_layernorm_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, void* y, int64_t ldy,
                       int64_t T, int creal, float out_scale, int dtype, void* stream);

/* Final reconstruction conv 3x3 (Cin stored channels -> 1) to an fp32 NCHW image, with add_mean folded
 * (out = out_scale * conv + out_bias).  w : [9][Cin].  Replaces tail[-1] + add_mean (:1303,1358). */
int rdst_last_conv_fwd(const void* x, int64_t ldx, const float* w, float bias, float out_scale, float out_bias,
                       float* img, int B, int H, int W, int Cin, int dtype, void* stream);

/* ---- backward kernels (fp32, CUDA cores): the training path of the module ------------------------------ */
/* Data gradients reuse the forward kernels: dX = dY.W is rdst_linear_fwd with the transposed weight; the conv
 * data gradient is rdst_conv3x3_fwd with the flipped/transposed filter.  The entry points below add the rest. */

/* dW[n][k] += sum_t dY[t][n] * X[t][k]  and, if db != NULL, db[n] += sum_t dY[t][n]   (fp32 atomics over token splits).
 * conv != 0: X rows are 3x3 neighbourhoods gathered from a [B*H*W][ldx] map (K = 9*Cin, Cin % 64 == 0), i.e. the
 * weight gradient of rdst_conv3x3_fwd.  Weight gradient of nn.Linear / nn.Conv2d on the path. */
int rdst_gemm_tn_acc(const float* dy, int64_t ldy, const float* x, int64_t ldx, float* dw, float* db, int64_t T,
                     int N, int K, int conv, int B, int H, int W, int Cin, void* stream);

/* Affine-free LayerNorm over `creal` real of K stored channels (dense_layout != 0: pads at [60,64) and the last two
 * of every 32-block after 64): y = (x-mean)*rstd, pads 0.  Backward: dx = rstd*(g - mean(g) - xhat*mean(g*xhat)) + resid + resid2
 * (both optional; dx may alias resid2 for in-place accumulation). */
int rdst_lnhat_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t T, int K, int creal, int dense_layout,
                   void* stream);
int rdst_lnhat_bwd(const float* dxhat, int64_t ldd, const float* x, int64_t ldx, const float* resid, int64_t ldr,
                   const float* resid2, int64_t ldr2, float* dx, int64_t ldo, int64_t T, int K, int creal,
                   int dense_layout, void* stream);

/* Backward of rdst_layernorm_fwd (affine, * scale): dx (may be NULL), dgamma += , dbeta += . */
int rdst_layernorm_bwd(const float* dy, int64_t ldd, const float* x, int64_t ldx, const float* gamma, float* dx,
                       int64_t ldo, float* dgamma, float* dbeta, int64_t T, int creal, float scale, void* stream);

/* y[t][n] += alpha * x[t][n] (gradient accumulation at residual joins). */
int rdst_axpy(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t T, int N, float alpha, void* stream);

/* Inverse of the pixel-shuffle store of rdst_conv3x3_fwd(shuffle=2): z[t][s*G+c] = u[(b,2y+dy,2x+dx)][c], s = 2dy+dx
 * (gradient of nn.PixelShuffle(2) in the token-major layout). */
int rdst_pixel_unshuffle2(const float* u, int64_t ldu, float* z, int64_t ldz, int B, int H, int W, int G, void* stream);

/* Exact-erf GELU on a [T][N] matrix, forward and backward (dx = dy * gelu'(x)). */
int rdst_gelu_fwd(const float* x, int64_t ldx, float* y, int64_t ldy, int64_t T, int N, void* stream);
int rdst_gelu_bwd(const float* x, int64_t ldx, const float* dy, int64_t ldd, float* dx, int64_t ldo, int64_t T, int N,
                  void* stream);

/* Backward of rdst_window_attention_fwd: dqkv [T][ldg] (q|k|v gradients, same layout as qkv) is overwritten,
 * dtable (225, heads) is accumulated. */
int rdst_window_attention_bwd(const float* qkv, int64_t ldq, const float* table, const float* dout, int64_t ldo,
                              float* dqkv, int64_t ldg, float* dtable, int B, int H, int W, int C, int heads, int shift,
                              void* stream);

/* ---- weight packing of the training path, forward and backward (rdst_b200/csrc/pack.cu) ------------------------------ */
/* Wp[pn(n)][pk(k)] = rs(n) * W[n][k] * gamma[k],  bp[pn(n)] = rs(n) * (b[n] + sum_k W[n][k]*beta[k])   (fp32)
 *   W [N][K], Wp [.][ldp] and bp pre-zeroed by the caller; gamma/beta = the LayerNorm in front of the Linear or NULL;
 *   rs(n) = q_scale for n < q_rows (attention scale on the q rows, swin_transformer_sr.py:120), else 1;
 *   pn / pk = stored position of a real channel in the padded dense-block layout if scatter_rows / scatter_cols, else
 *   identity.  Same arithmetic as rdst_b200/packing.py:pack_stl / pack_dstl_tail. */
int rdst_pack_linear_fwd(const float* W, const float* b, const float* gamma, const float* beta, float* Wp, float* bp,
                         int N, int K, int ldp, int scatter_rows, int scatter_cols, int q_rows, float q_scale, void* stream);
/* Gradients of the parameters from the gradients of the packed tensors: dW, db are overwritten, dgamma / dbeta are
 * accumulated (pass them zeroed). */
int rdst_pack_linear_bwd(const float* W, const float* gamma, const float* beta, const float* dWp, const float* dbp,
                         float* dW, float* db, float* dgamma, float* dbeta, int N, int K, int ldp, int scatter_rows,
                         int scatter_cols, int q_rows, float q_scale, void* stream);

/* Batched form of the two entry points above: all Linears of one RDSTB (<= RDST_PACK_MAX) in ONE launch; the descriptors
 * are copied into the kernel's parameter space, nothing is read from `descs` after the call returns.
 * backward == 0: fields W, b, gamma, beta, Wp, bp are used; backward != 0: W, gamma, beta, dWp, dbp, dW, db, dgamma, dbeta. */
#define RDST_PACK_MAX 32
typedef struct RdstPackDesc {
  const float* W; const float* b; const float* gamma; const float* beta;
  float* Wp; float* bp;
  const float* dWp; const float* dbp;
  float* dW; float* db; float* dgamma; float* dbeta;
  int N, K, ldp, scatter_rows, scatter_cols, q_rows;
  float q_scale;
  int _pad;
} RdstPackDesc;
int rdst_pack_linear_batch(const RdstPackDesc* descs, int n, int backward, void* stream);

/* ---- tensor-core GEMMs of the training path (precision 'bf16'): fp32 storage, operands rounded to bf16 while they
 *      are staged into shared memory, tcgen05.mma with fp32 accumulation in TMEM (rdst_b200/csrc/tc_train.cu) ------ */

/* Y[t][n] = out_scale * ( op(X[t][0:K]) . Wop + bias[n] ) * G[t][n] + R[t][n]          (all fp32 in memory)
 *   a_op: 0 none; 1 LNhat over ln_creal real channels (as rdst_linear_fwd); 2 exact-erf GELU (fc2 reads the saved
 *         pre-activation of fc1, the activated tensor is never stored).
 *   G = gelu'(gelu_aux[t][n]) if gelu_aux != NULL (GELU backward fused into the fc2 data gradient), else 1.
 *   w_mn_major == 0: W is [N][ldw], K contiguous   (forward nn.Linear weight; conv filter [N][9][Cin])
 *   w_mn_major == 1: W is [K][ldw], N contiguous   (the forward weight of a Linear used for its data gradient dX = dY.W)
 *   conv != 0: X is a [B*H*W][ldx] map, K = 9*Cin, the A rows are 3x3 neighbourhoods (zero padding); shuffle == 2 folds
 *   PixelShuffle(2) into the store (N == 256 = 4 sub-pixels x 64 channels, Y is the [B*2H*2W][ldy] map).
 *   Activation rows must be 16-byte aligned and readable up to the next multiple of 8 columns (pads may hold anything).
 *   Same layers as rdst_linear_fwd / rdst_conv3x3_fwd / rdst_gelu_fwd / rdst_gelu_bwd. */
int rdst_gemm_tc(const float* x, int64_t ldx, const float* w, int64_t ldw, int w_mn_major, const float* bias,
                 const float* resid, int64_t ldr, const float* gelu_aux, int64_t lda, float* y, int64_t ldy,
                 int64_t T, int K, int N, int a_op, int ln_creal, float out_scale, int conv, int B, int H, int W,
                 int Cin, int shuffle, void* stream);

/* Data gradient of LayerNorm-hat -> Linear in one kernel:  dxh = scale * dY[T][K] . W[K][N]  (W = forward weight, [out=K][in=N]),
 * then, in the epilogue, the LayerNorm-hat backward of every row with respect to X[t][0:N] over `creal` real channels
 * (pads of the dense-block layout excluded, pad outputs 0) plus up to two residual rows:
 *   dX = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)) + R + R2.      N <= 128, N % 16 == 0; dX may alias R2.
 * = rdst_gemm_tc(w_mn_major = 1) followed by rdst_lnhat_bwd(dense_layout = 1), without storing dxh. */
int rdst_gemm_tc_lnbwd(const float* dy, int64_t ldy, const float* w, int64_t ldw, const float* x, int64_t ldx,
                       const float* resid, int64_t ldr, const float* resid2, int64_t ldr2, float* dx, int64_t ldo,
                       int64_t T, int K, int N, int creal, float scale, void* stream);

/* Tensor-core version of rdst_gemm_tn_acc: dW[n][k] += sum_t dY[t][n] * op(X)[t][k], db[n] += sum_t dY[t][n].
 * x_op: 0 none, 1 LNhat over x_creal real channels (K <= 240), 2 exact-erf GELU -- the normalised / activated operand of
 * the weight gradient is recomputed while staging instead of being stored.  Cin % 16 == 0 in conv mode. */
int rdst_gemm_tn_tc(const float* dy, int64_t ldy, const float* x, int64_t ldx, float* dw, float* db, int64_t T,
                    int N, int K, int conv, int B, int H, int W, int Cin, int x_op, int x_creal, void* stream);

/* Window attention of the training path on tcgen05 (rdst_b200/csrc/tc_attn_train.cu): same contract as
 * rdst_window_attention_fwd / _bwd in fp32 storage (qkv [T][ldq] = q|k|v with q pre-scaled, table (225, 6), heads = 6,
 * C in {60, 90, 120}), operands rounded to bf16 while staged.  The forward also writes the row log-sum-exp
 * lse [T][6] (may be NULL), which the backward reads instead of recomputing the softmax reduction.
 * Replaces WindowAttention.forward (swin_transformer_sr.py:110-141) + roll / window_partition / window_reverse / mask
 * (:32-59, :211-232, :239-271) and their autograd. */
int rdst_window_attention_tc_fwd(const float* qkv, int64_t ldq, const float* table, float* out, int64_t ldo, float* lse,
                                 int B, int H, int W, int C, int shift, void* stream);
int rdst_window_attention_tc_bwd(const float* qkv, int64_t ldq, const float* table, const float* lse, const float* dout,
                                 int64_t ldo, float* dqkv, int64_t ldg, float* dtable, int B, int H, int W, int C, int shift,
                                 void* stream);

/* ---- tcgen05 / TMEM kernels (bf16 operands, fp32 accumulate), sm_100a only ---------------------------- */

/* Fused Swin MLP:  Y[t] = X[t] + fc2( GELU( fc1( LNhat(X[t]) ) ) ),  bf16 storage, T tokens, C in {60,90,120}
 * (stored width Cp = 64/96/128, hidden 2C padded to Hp = 128/192/240).  LNhat as in rdst_linear_fwd.
 *   w1img: fc1 weight (gamma folded) as a ready-made UMMA K-major operand image  [Cp/8][Hp][8] fp16
 *   w2img: HALF the fc2 weight (the GELU stage emits 2*GELU; exact scaling)    [Hp/8][Cp][8] fp16 (hidden is fp16)
 *   b1 [Hp] (beta folded), b2 [Cp] fp32.  exact_gelu != 0 evaluates erff instead of the fitted tanh form.
 * Replaces norm2 + Mlp + residual, swin_transformer_sr.py:272 and :23-29.  X and Y may not alias. */
int rdst_stl_mlp_fwd_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, const void* w1img, const void* w2img,
                          const float* b1, const float* b2, int64_t T, int C, int exact_gelu, void* stream);

/* Same MLP, fused with the DenseSTLayer tail: the block output y is not stored; instead
 *   dense[t][0:32] = dense_scale * ( LNhat(y[t]) . Wt^T + bt )     (30 real growth channels + 2 zero pads)
 * is written into the dense buffer slice (`dense` points at &D[0][64+32j], row stride ldd) -- the reference's
 * tail LayerNorm + Linear + mul + torch.cat (rdst_variations.py:339-340) without materialising y or the concat.
 *   wtimg : [Cp/8][32][8] bf16 K-major image of the tail Linear (LN gamma folded), bt [32] fp32 (beta folded). */
int rdst_stl_mlp_tail_fwd_bf16(const void* x, int64_t ldx, const void* w1img, const void* w2img, const float* b1,
                               const float* b2, const void* wtimg, const float* bt, void* dense, int64_t ldd,
                               float dense_scale, int64_t T, int C, int exact_gelu, void* stream);

/* Fused shifted-window attention block:
 *   Y[t] = X[t] + proj( softmax_j( q_i.k_j + bias(i,j) + mask(i,j) ) v_j ),  [q|k|v] = qkv(LNhat(X)), bf16 storage,
 * over 8x8 windows of a (B,H,W) token grid with cyclic shift `shift` in {0,4}; C in {60,90,120}, 6 heads.
 * The roll/window_partition/window_reverse/roll of the reference are index arithmetic in the gather/scatter; the
 * {0,-100} mask is evaluated in closed form; scores/probabilities live only in TMEM/shared memory.
 *   wqkv_img : 6 per-head K-major operand images [Cp/8][NH][8] bf16, rows = q_h | k_h | v_h | zero pad
 *              (NH = 32/48/64), LN gamma, head_dim^-0.5 and log2(e) folded into the q rows
 *   bqkv     : [6][NH] fp32 (LN beta folded);  wproj_img : [KPROJ/8][Cp][8] bf16 with K index = head*HDO + d
 *              (HDO = 16/16/20, KPROJ = 96/96/128);  bproj : [Cp];
 *   table    : [6][15][24] 32-bit words, each a packed fp16 pair (t[h][dy+7][dx+7], t[h][dy+7][dx+6]) with
 *              t = relative_position_bias_table[(dy+7)*15+(dx+7)][h] * log2(e)  (the softmax runs on packed fp16
 *              pairs; row pitch 24 keeps the per-row lookups free of shared-memory bank conflicts)
 * Replaces swin_transformer_sr.py:239-271 (block) and :110-141 (WindowAttention).  X and Y must not alias. */
int rdst_stl_attn_fwd_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, const void* wqkv_img,
                           const void* wproj_img, const float* bqkv, const float* bproj, const float* table,
                           int B, int H, int W, int C, int shift, void* stream);

/* 3x3 convolution as an implicit GEMM on tcgen05 (bf16 storage).  Same contract as rdst_conv3x3_fwd, for the
 * shapes of the network: (Cin,N) = (160,64) LFF, (64,64) conv_after_body, (64,256)+shuffle=2 up-sampler.
 *   wimg : N/NT slices (NT = 32 when Cin == 160, 128 when N == 256, else 64), each 9 taps of a K-major operand image [Cin/8][NT][8] bf16
 *          built from the [N][9][Cin] weight of rdst_conv3x3_fwd (tap = ky*3+kx).
 * The halo tile is staged once in shared memory; the nine taps are descriptor offsets into that image.
 * (160,64) runs on CTA pairs (tcgen05.mma.cta_group::2, M = 256): the two 32-column slices of wimg are the two
 * CTAs' halves of the B operand.  y must be 32-byte aligned with ldy a multiple of 16 elements (256-bit stores). */
int rdst_conv3x3_fwd_bf16_tc(const void* x, int64_t ldx, const void* wimg, const float* bias, const void* resid,
                             int64_t ldr, void* y, int64_t ldy, int B, int H, int W, int Cin, int N,
                             float out_scale, int shuffle, void* stream);

/* Final reconstruction conv (64 stored channels -> 1) on tcgen05 (N padded to 16), fp32 NCHW image out,
 * add_mean folded: img = (conv + bias) * out_scale + out_bias.
 *   wimg : one K-major operand image [8][16][8] bf16 whose rows are the 9 filter taps (row t = w[t][0:64], rows 9..15
 *          zero): the kernel computes the per-position tap products on the tensor core and the 9-point sum on CUDA cores.
 * Same contract as rdst_last_conv_fwd with Cin = 64.  Replaces tail[-1] + add_mean (rdst_variations.py:1303,1358). */
int rdst_last_conv_fwd_bf16_tc(const void* x, int64_t ldx, const void* wimg, float bias, float out_scale,
                               float out_bias, float* img, int B, int H, int W, void* stream);

/* ---- LR synthesis and image metrics (rdst_b200/csrc/imaging.cu) -------------------------------------- */
/* cv2.resize(img, dsize, interpolation=cv2.INTER_CUBIC) of B single-channel fp32 images [B][Hs][Ws] -> [B][Hd][Wd], as
 * MedicalImageBasicDataset.resize uses it for LR synthesis and the bicubic "res" images (datasets/basic_dataset.py:65-123,
 * :258-301).  ix/cx ([Wd][4]) and iy/cy ([Hd][4]) are the clamped tap indices and cubic weights (A = -0.75) of every output
 * column / row, built by the host (rdst_b200/imaging.py) exactly as OpenCV derives them; 16-byte aligned. */
int rdst_bicubic_resize_f32(const float* src, float* dst, const int* ix, const float* cx, const int* iy, const float* cy,
                            int B, int Hs, int Ws, int Hd, int Wd, void* stream);
/* out[b] += sum_i (a[b][i] - b[b][i])^2 in fp64 (out pre-zeroed): MSE / PSNR of metrics/sr_metrics.py:8-9. */
int rdst_sqdiff_sum_f64(const float* a, const float* b, double* out_zeroed, int B, int64_t n_per_image, void* stream);
/* out[b] += sum of the SSIM map over the valid region [3,H-3) x [3,W-3) (7x7 uniform window, sample covariance, K1 = 0.01,
 * K2 = 0.03, data_range 1): skimage.metrics.structural_similarity as metrics/sr_metrics.py:12-13 calls it; the caller
 * divides by (H-6)(W-6). */
int rdst_ssim_sum_f64(const float* a, const float* b, double* out_zeroed, int B, int H, int W, void* stream);

/* Debug hook: when given a device buffer of 128 uint64, CTA 0 of every following rdst_stl_attn_fwd_bf16 launch records
 * clock64() at its phase boundaries (64 stamps per warpgroup).  Pass NULL to switch it off (the default). */
int rdst_debug_attn_timing(void* device_buffer_128_u64);
/* Selects the kernel behind rdst_stl_attn_fwd_bf16: 2 (default) = warp-specialised pipeline (csrc/tc_attn2.cu),
 * 1 = the round-1 lock-step kernel (csrc/tc_attn.cu), kept for A/B timing.  Same arguments, same contract. */
int rdst_debug_attn_variant(int variant);
/* Role timelines of CTA 0 of the warp-specialised attention kernel: 5 roles x 256 uint64 (slot 0 unused, then
 * clock64() stamps in program order of one thread of the role).  NULL switches it off. */
int rdst_debug_attn2_timing(void* device_buffer_1280_u64);
int rdst_debug_mlp_timing(void* device_buffer_128_u64);     /* same for rdst_stl_mlp_*_fwd_bf16 */
/* Kernel behind rdst_stl_mlp_fwd_bf16 / rdst_stl_mlp_tail_fwd_bf16: 2 (default) = warp-specialised pipeline (csrc/tc_mlp2.cu),
 * 1 = the round-1 lock-step kernel (csrc/tc_mlp.cu), kept for A/B timing; rdst_debug_mlp2_timing: role timelines of CTA 0 of the
 * warp-specialised kernel (5 x 256 uint64, like rdst_debug_attn2_timing). */
int rdst_debug_mlp_variant(int variant);
int rdst_debug_mlp2_timing(void* device_buffer_1280_u64);
int rdst_debug_conv_timing(void* device_buffer_128_u64);    /* same for rdst_conv3x3_fwd_bf16_tc (128 stamps, CTA 0) */

/* tcgen05 issue-pattern microbenchmark (timing only, zero operands): `count` MMAs M=128 x N x K=16 issued round-robin over
 * `chains` accumulators by one lane, then one commit.  out[0] = cycles first issue -> completion, out[1] = issue cycles.
 * a_tmem: A from TMEM (else shared memory); masked: disable-output-lane form.  Used by tools/umma_bench.py only. */
int rdst_umma_bench(int N, int chains, int count, int a_tmem, int masked, void* out_2_u64, void* stream);

/* TMEM <-> register bandwidth probe (one CTA): nwarps warps each move 16 KB per round with four tcgen05.ld.32x32b.x32
 * (store != 0: tcgen05.st).  out[0] = cycles of the slowest warp, out[1] = bytes moved.  tools/tmem_bw.py only. */
int rdst_tmem_bw_bench(int nwarps, int reps, int store, void* out_3_u64, void* stream);

/* Self-test of the UMMA plumbing: D[M=128][N] = A[128][K] . B[N][K]^T with bf16 inputs, fp32 output.
 * b_mn_major != 0 feeds B from an MN-major shared-memory image.  Used by tests/ only. */
int rdst_umma_selftest(const void* a_bf16, const void* b_bf16, float* d, int N, int K, int b_mn_major,
                       int m64, void* stream);

/* Self-test of the TMA plumbing: loads the 8x8 window whose shifted-frame origin is (hs0, ws0) of image b from the
 * token-major activation x ([B][H][W][ldx] bf16, C channels) as four 4x4-token boxes per 64-channel panel
 * (SWIZZLE_128B), dumps the raw shared-memory image (ceil(C/64) * 8192 bytes) and stores the tile to the same
 * window of y.  Used by tests/ only. */
int rdst_tma_selftest(const void* x, int64_t ldx, void* y, int64_t ldy, int B, int H, int W, int C, int shift, int b,
                      int hs0, int ws0, void* dump, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RDST_B200_H */
