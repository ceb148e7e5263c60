/*
 * diffews_b200 — C ABI of the B200 (sm_100a) hot path of DiffewS.
 *
 * Every entry point is `extern "C"`, takes raw DEVICE pointers + sizes + a cudaStream_t (as void*), returns an int
 * status (0 = DFW_OK, negative = error), never throws, never allocates device memory, reads no environment variable
 * and is stream-ordered.  The only process-wide state is the explicit option table below (dfw_set_option): kernel-
 * selection switches for A/B measurements, every default being what bench.py measures.  There is NO CPU / other-arch fallback: on a non-sm_100 device every compute entry point
 * returns DFW_ERR_ARCH.
 *
 * The reference (ga1i13o/DiffewS) has no native code; each entry point below replaces a *library call site* of the
 * reference's PyTorch path.  "ref:" comments cite /root/reference file:line; "(upstream)" marks diffusers-0.25
 * modules that the reference imports (requirements.txt:2) and whose source is not vendored in the reference.
 *
 * Activation layout everywhere: channels-last (NHWC) — a [N,H,W,C] image tensor is the same memory as the
 * [N, H*W, C] token tensor the transformer blocks use.  bf16 storage, fp32 accumulation / statistics.
 */
#ifndef DIFFEWS_B200_H_
#define DIFFEWS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFW_OK 0
#define DFW_ERR_INVALID (-1) /* bad argument (shape / alignment / unsupported combination) */
#define DFW_ERR_CUDA (-2)    /* a CUDA runtime/driver call failed (message on stderr)        */
#define DFW_ERR_ARCH (-3)    /* device is not sm_100 (B200) — no fallback by design           */

/* Process-wide options (dfw_set_option / dfw_get_option).  Defaults in brackets; every option only selects between
 * kernels that compute the same result. */
#define DFW_OPT_PDL 0               /* [0] programmatic dependent launch on the hot kernels (correct, neutral on B200)   */
#define DFW_OPT_T128 1              /* [1] channel-major conv kernel igemm_t128 (0: everything on igemm_kernel)          */
#define DFW_OPT_T128_MAXC 2         /* [1<<20] largest Cout routed to igemm_t128                                          */
#define DFW_OPT_HALO 3              /* [1] halo mainloop of the generic 3x3 kernel (0: one TMA box per filter tap)       */
#define DFW_OPT_GN_CTAS_PER_SM 4    /* [3] GroupNorm grid sizing                                                          */
#define DFW_OPT_PREPROC_TWO_PASS 5  /* [0] data layer: two-launch resample with a uint8 intermediate in HBM              */
#define DFW_OPT_ATTN_V2 6           /* [0] round-1 attention kernel (P through smem, two passes over S)                  */
#define DFW_OPT_SEG_HEAD 7          /* [1] fused decoder head dfw_seg_head_u8 (0: gn-apply + conv 128->3 + seg_post)     */
#define DFW_OPT_ATTN_BWD_UNFUSED 8  /* [0] round-1 attention backward (batched GEMMs over materialised logits)            */
#define DFW_OPT_ATTN_V4 9           /* [0] attention v4: 96-key tiles, private S double buffers + one MMA issuer per query tile */
#define DFW_OPT_B_RESIDENT 10       /* [0] weight-stationary mainloop for 1x1 / linear layers with a short K (bit-identical, neutral) */
#define DFW_OPT_COUNT 11
int dfw_set_option(int option, int value); /* DFW_ERR_INVALID for an unknown option */
int dfw_get_option(int option);            /* current value; -1 for an unknown option */

/* epilogue flags for dfw_conv2d_igemm / dfw_linear */
#define DFW_EPI_OUT_F32 1   /* y is fp32 (default: 16-bit, bf16 or fp16 per DFW_EPI_F16)              */
#define DFW_EPI_RES_F32 2   /* residual is fp32 (default: 16-bit)                                     */
#define DFW_EPI_GEGLU 4     /* weight rows are [128 value | 128 gate] interleaved per 256-row block;  */
                            /* y[:, j] = (v+bv) * gelu_erf(g+bg), y has Cout/2 channels               */
#define DFW_EPI_SILU 8      /* y = silu(acc + bias) (time-embedding MLP)                              */
#define DFW_EPI_F16 16      /* every 16-bit tensor of the call (x, w, 16-bit y / residual) is IEEE fp16 instead of  */
                            /* bf16.  tcgen05 kind::f16 needs A and B in the SAME format (mixing traps on sm_100).  */

/* ABI version: bump on any signature change. */
int dfw_version(void);
/* DFW_OK when the current CUDA device is a B200-class (sm_100) GPU. */
int dfw_device_ok(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches claim). */
long long dfw_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * K3/K6  tcgen05 implicit-GEMM convolution and Linear (TMA-fed, TMEM accumulators, persistent CTAs).
 * ref: every nn.Conv2d / nn.Linear reached from diffews/models/unet_2d_condition.py:1118-1121,1161,1191,1226,1249
 *      (ResnetBlock2D / Downsample2D / Upsample2D / Transformer2DModel / BasicTransformerBlock, upstream) and from
 *      diffews/marigold_pipeline_rgb_latent_noise.py:852-853,901-902 (AutoencoderKL encoder/decoder, upstream).
 *
 *   x        bf16 [N, Hin, Win, Cin]                Cin % 64 == 0
 *   w        bf16 [Cout, ksize, ksize, Cin]         (= torch weight.permute(0,2,3,1)), K index = (kh*ks+kw)*Cin + c
 *   bias     fp32 [Cout] or, if bias_sample_stride != 0, [N, bias_sample_stride] (per-image bias: conv bias +
 *            time-embedding projection, ResnetBlock2D temb add (upstream)); may be NULL
 *   residual bf16/fp32 [N, Hout, Wout, Cout_eff] or NULL;   y = (acc + bias) * out_scale + residual
 *   y        bf16/fp32 [N, Hout, Wout, Cout_eff]            Cout_eff = Cout (Cout/2 with DFW_EPI_GEGLU)
 *   ksize 1|3; stride 1|2; pad_mode 0: symmetric (ksize-1)/2 ; 1: VAE-encoder downsample F.pad(0,1,0,1) + pad 0.
 * ------------------------------------------------------------------------------------------------------------ */
int dfw_conv2d_igemm(const void* x, const void* w, const float* bias, int bias_sample_stride, const void* residual,
                     void* y, int N, int Hin, int Win, int Cin, int Cout, int ksize, int stride, int pad_mode,
                     int flags, float out_scale, void* stream);

/* y[M, Nout_eff] = epi(x[M,K] @ w[Nout,K]^T + bias).  K % 64 == 0.  ref: nn.Linear in Attention.to_q/k/v/to_out,
 * Transformer2DModel.proj_in/out, FeedForward (GEGLU + Linear), TimestepEmbedding (upstream). */
int dfw_linear(const void* x, const void* w, const float* bias, const void* residual, void* y, int M, int K,
               int Nout, int flags, float out_scale, void* stream);

/* Nearest-2x upsample fused with the 3x3 / pad 1 convolution that follows it (diffusers Upsample2D: F.interpolate(
 * scale_factor=2, "nearest") -> conv, upstream; UNet up blocks unet_2d_condition.py:1226, VAE decoder pipeline:901).
 * Each output phase (oh%2, ow%2) only ever sees 2x2 distinct input pixels, so the 9 taps collapse into 4 with
 * pre-summed weights: 2.25x fewer FLOPs and the 4x larger upsampled tensor is never written.
 *   x  16-bit [N, Hin, Win, Cin];  y 16-bit / fp32 [N, 2Hin, 2Win, Cout]  (Cout*elem % 16 == 0)
 *   w4 16-bit [4 phases (ph*2+pw)][Cout][(a*2+b)*Cin + c] with, per axis, phase 0: {W[0], W[1]+W[2]},
 *      phase 1: {W[0]+W[1], W[2]} (see diffews_b200/weights.py::upconv_phase_weights). */
int dfw_upconv2x_igemm(const void* x, const void* w4, const float* bias, void* y, int N, int Hin, int Win, int Cin,
                       int Cout, int flags, float* gn_partial /* NULL or 4 * dfw_gn_partial_floats(N) floats */,
                       void* stream);

/* Convolution that also emits the GroupNorm(32) statistics of its OUTPUT (the tensor the next layer normalises), so
 * the consumer can skip its statistics pass (one full read of the tensor): every epilogue thread reduces its row's
 * channel groups, warps reduce-scatter with shuffles, and each CTA writes one partial (sum, sum of squares) per
 * (image, group): gn_partial[N][2 x #SMs][32][2] fp32 (one slot per epilogue warp group), dfw_gn_partial_floats(N) floats.  Requires Cout/32 in {4, 8, 16},
 * >= 128 output pixels per image row block (one image per tile) and the TMA epilogue (Cout*elem % 16 == 0, residual
 * of the output's element size).  Consumed by dfw_groupnorm_from_partial(nchunks = dfw_gn_partial_floats(N)/(N*64)).
 * ref: the GroupNorm at the head of every diffusers ResnetBlock2D / Attention block of the VAE (upstream). */
long long dfw_gn_partial_floats(int N);
int dfw_conv_gnstats_supported(int N, int Hout, int Wout, int Cout);   /* 1 if the fused statistics apply */
int dfw_conv2d_igemm_gnstats(const void* x, const void* w, const float* bias, const void* residual, void* y, int N,
                             int Hin, int Win, int Cin, int Cout, int ksize, int stride, int pad_mode, int flags,
                             float out_scale, float* gn_partial, void* stream);

/* Batched GEMM with per-batch "weights": y[b] (M x Nout) = epi(x[b] (M x K) @ w[b]^T + bias), w[b] element (n, k) at
 * w[b*w_batch_stride + n*w_row_stride + k] (strides in elements, multiples of 8).  M >= 128, K % 64 == 0.
 * ref: the single-head d=512 attention of the VAE mid block (diffusers Attention + AttnProcessor2_0, upstream),
 *      reached from pipeline:852 / :901 — Q K^T (fp32 logits) and P V per image. */
int dfw_bmm_nt(const void* x, const void* w, long long w_row_stride, long long w_batch_stride, const float* bias,
               void* y, int B, int M, int K, int Nout, int flags, float out_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K1  KV-fused flash attention forward, head_dim 64, tcgen05 + TMEM + TMA.
 * ref: diffews/models/attention_processor.py:251-271 (MyXFormersAttnProcessor: key = cat([key_self, bank_folded]),
 *      xformers.ops.memory_efficient_attention(q,k,v,scale)); :351-365 (SDPA variant).
 *  The concatenation is never materialised: keys/values stream from two sources (self, then bank).
 *   (16-bit tensors are bf16, or fp16 when f16 != 0 — one format per call)
 *   q       16-bit, element (b, l, h, d) at q[b*q_batch_stride + l*q_row_stride + h*64 + d],   l < Lq
 *   k_self / v_self  same addressing with kv_self_* strides, l < Ls
 *   k_bank / v_bank  same addressing with kv_bank_* strides, l < Lb (Lb = k_shots * S, shot-major; may be 0/NULL)
 *   o       16-bit [B, Lq, heads*64] (row stride o_row_stride)
 *  softmax(q k^T * scale) v over the Ls + Lb keys, fp32 softmax, fp32 accumulation.
 * ------------------------------------------------------------------------------------------------------------ */
int dfw_attn_kvfused_fwd(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                         const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                         const void* k_bank, const void* v_bank, long long kv_bank_batch_stride,
                         int kv_bank_row_stride, void* o, long long o_batch_stride, int o_row_stride, int B,
                         int heads, int Lq, int Ls, int Lb, float scale, int f16, void* stream);

/* K1 with the softmax statistics the backward needs: lse fp32 [B, heads, Lq] = log2-domain logsumexp of the scaled logits
 * (log2(sum_j 2^(s_j * scale * log2 e))).  Same arguments and result as dfw_attn_kvfused_fwd otherwise. */
int dfw_attn_kvfused_fwd_lse(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                             const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                             const void* k_bank, const void* v_bank, long long kv_bank_batch_stride,
                             int kv_bank_row_stride, void* o, long long o_batch_stride, int o_row_stride, int B,
                             int heads, int Lq, int Ls, int Lb, float scale, int f16, float* lse, void* stream);

/* K1b  BACKWARD of K1 (BASELINE config 4: training-shape forward + backward of the KV-fused attention; ref: autograd of
 *      xformers.ops.memory_efficient_attention on cat([key, folded bank]) -- attention_processor.py:251-271 under
 *      train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1374-1391).  Same tensor conventions as the forward;
 *      o = the forward output, d_o its gradient (same strides as o); dq like q, dk/dv_self like k/v_self, dk/dv_bank like
 *      the bank; lse = the statistics of dfw_attn_kvfused_fwd_lse, or NULL (they are then recomputed by one forward into
 *      the workspace).  Flash-style and fused: the L_q x L_k matrices exist only as 128 x 128 tiles in tensor memory
 *      (csrc/attn_bwd_fused.cu: a dK/dV kernel over key tiles, a dQ kernel over query tiles x key slices, all products on
 *      tcgen05); deterministic (no atomics); any Lq / Ls / Lb >= 1.  Workspace (dfw_attn_bwd_workspace_bytes, 256-B aligned):
 *      O(B * Lq * heads * 64) bytes.  DFW_OPT_ATTN_BWD_UNFUSED selects the round-1 path (logits materialised in the
 *      workspace, L multiples of 64) for A/B measurements. */
long long dfw_attn_bwd_workspace_bytes(int B, int heads, int Lq, int Ls, int Lb);
int dfw_attn_kvfused_bwd(const void* q, long long q_batch_stride, int q_row_stride, const void* k_self,
                         const void* v_self, long long kv_self_batch_stride, int kv_self_row_stride,
                         const void* k_bank, const void* v_bank, long long kv_bank_batch_stride, int kv_bank_row_stride,
                         const void* o, const void* d_o, long long o_batch_stride, int o_row_stride, const float* lse,
                         void* dq, void* dk_self, void* dv_self, void* dk_bank, void* dv_bank, int B, int heads, int Lq, int Ls,
                         int Lb, float scale, int f16, void* workspace, void* stream);

/* K2  cross-attention to a short prompt embedding (Lctx <= 128 keys, head_dim 64), CUDA cores.
 * ref: BasicTransformerBlock.attn2 (upstream) reached from unet_2d_condition.py:1161; Lctx = 2 at eval
 *      (marigold_pipeline_rgb_latent_noise.py:591-601).
 *   q 16-bit [B, L, heads*64]; k,v [B, Lctx, heads*64] (kv_batch_stride elements, 0 = shared); o like q;
 *   f16 != 0: fp16 tensors, else bf16. */
int dfw_cross_attn_fwd(const void* q, const void* k, const void* v, long long kv_batch_stride, void* o, int B,
                       int L, int heads, int Lctx, float scale, int f16, void* stream);

/* K2b  the same attn2 block for a FIXED short prompt (one empty-prompt embedding shared by every sample, Lctx = 2 at
 * eval: marigold_pipeline_rgb_latent_noise.py:591-601, :690-692), collapsed: with K, V constant,
 *   to_out(softmax(to_q(x) K^T scale) V) + b  ==  b + sum_{h,j} softmax_j(x @ Wlog^T)[h,j] * U[(h,j)],
 *   Wlog[(h,j)] = scale K[j,h] @ Wq[h]  ([heads*Lctx, C], applied with dfw_linear -> `logits`),  U[(h,j)] = Wo[:,h] @ V[j,h].
 * logits fp32 [M, ld_logits] (first heads*Lctx columns used, <= 96); U fp32 [heads*Lctx, C]; bias fp32 [C];
 * residual (nullable) / out [M, C] of `dtype` (0 bf16, 1 fp32, 2 fp16). */
int dfw_cross_attn_collapsed(const float* logits, int ld_logits, const float* U, const float* bias, const void* residual,
                             void* out, int dtype, long long M, int C, int heads, int Lctx, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K4  GroupNorm(32 groups) (+SiLU), NHWC, fp32 statistics, deterministic two-stage reduction.
 * ref: nn.GroupNorm(+SiLU) in ResnetBlock2D / Transformer2DModel.norm / conv_norm_out
 *      (diffews/models/unet_2d_condition.py:1246-1248; upstream blocks).
 *   x [N, HW, C] with x_dtype 0 = bf16, 1 = fp32, 2 = fp16; gamma/beta fp32 [C]; y bf16 (fp16 if y_f16: normalised activations are
 *   bounded, and fp16 carries 3 more mantissa bits into the tensor core) [N, HW, C].  C % 8 == 0, C % groups == 0.
 *   workspace: fp32, dfw_groupnorm_workspace_bytes(N, HW, C) bytes.
 * ------------------------------------------------------------------------------------------------------------ */
long long dfw_groupnorm_workspace_bytes(int N, int HW, int C, int groups);
int dfw_groupnorm_silu(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_f16, int N,
                       int HW, int C, int groups, float eps, int apply_silu, void* workspace, void* stream);

/* GroupNorm (+SiLU) whose statistics were already produced as partial sums (by dfw_conv2d_igemm_gnstats / dfw_upconv2x_igemm
 * or by any producer with the layout partial[N][nchunks][groups][2]): only the normalise-and-store pass runs. */
int dfw_groupnorm_from_partial(const void* x, int x_dtype, const float* partial, int nchunks, const float* gamma,
                               const float* beta, void* y, int y_f16, int N, int HW, int C, int groups, float eps,
                               int apply_silu, void* stream);

/* GroupNorm + SiLU folded into the consuming convolution (diffusers ResnetBlock2D: norm -> nonlinearity -> conv, upstream,
 * reached from diffews/marigold_pipeline_rgb_latent_noise.py:852-853,901-902): dfw_gn_scale_shift turns the statistics a
 * producing conv emitted (gn_partial, nchunks as above) into per-(image, channel) scale / shift, fp32 [N][2][C];
 * dfw_conv2d_igemm_gnin computes conv(silu(x * scale + shift)) + bias (+ residual) without materialising the normalised
 * tensor in HBM (x, w, y, residual 16-bit of one format, flags = DFW_EPI_F16 or 0; stride 1, pad (ksize-1)/2; optional
 * statistics of ITS output in gn_partial_out).  scratch: dfw_conv_gnin_scratch_bytes() bytes of device memory, 128-byte
 * aligned, contents irrelevant, owned by the caller and not shared between concurrently running calls (the kernel keeps a
 * small per-CTA ring of transformed tiles there; it stays in L2).  Only for the shapes dfw_conv_gnin_supported accepts
 * (Cout % 128 == 0, H, W % 16 == 0, enough tiles to fill the GPU); otherwise use dfw_groupnorm_from_partial +
 * dfw_conv2d_igemm. */
/* GroupNorm (+SiLU) BACKWARD (BASELINE config 4; the nn.GroupNorm -> SiLU autograd of every diffusers ResnetBlock2D in the
 * training step train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1320-1396).  x, dy, dx [N, HW, C] of one dtype
 * (0 bf16, 1 fp32, 2 fp16); dgamma, dbeta fp32 [C]; statistics are recomputed from x; deterministic (fixed-order folds). */
long long dfw_groupnorm_bwd_workspace_bytes(int N, int HW, int C, int groups);
int dfw_groupnorm_silu_bwd(const void* x, const void* dy, int dtype, const float* gamma, const float* beta, void* dx,
                           float* dgamma, float* dbeta, int N, int HW, int C, int groups, float eps, int apply_silu,
                           void* workspace, void* stream);

int dfw_gn_scale_shift(const float* partial, int nchunks, const float* gamma, const float* beta, float* scale_shift,
                       int N, long long HW, int C, int groups, float eps, void* stream);
int dfw_conv_gnin_supported(int N, int H, int W, int Cin, int Cout, int ksize);
/* 1 when a plain stride-1 16-bit convolution / residual projection of this shape is routed to the channel-major kernel
 * (igemm_t128_kernel) by dfw_conv2d_igemm / dfw_linear: what bench.py uses to attribute launches to that kernel. */
int dfw_conv_t128_eligible(int N, int H, int W, int Cin, int Cout, int ksize);
int dfw_conv2d_igemm_gnin(const void* x, const float* gn_scale_shift, const void* w, const float* bias,
                          const void* residual, void* y, int N, int Hin, int Win, int Cin, int Cout, int ksize, int flags,
                          float* gn_partial_out, void* scratch, void* stream);
long long dfw_conv_gnin_scratch_bytes(void);

/* K7  LayerNorm over the last dim. x [M, C] (x_dtype 0 bf16 / 1 fp32 / 2 fp16) -> y bf16 (fp16 if y_f16) [M, C].
 * C % 8 == 0, C <= 2048.
 * ref: BasicTransformerBlock.norm1/2/3 (upstream). */
/* LayerNorm backward (statistics recomputed from x): x, dy, dx [M, C] of one dtype (0 bf16 / 1 fp32 / 2 fp16);
 * dgamma, dbeta fp32 [C]; C % 8 == 0, C <= 1280; workspace: dfw_layernorm_bwd_workspace_bytes(M, C).  Deterministic.
 * ref: autograd of BasicTransformerBlock.norm1/2/3 in the training step (train...v3.py:1386 accelerator.backward). */
long long dfw_layernorm_bwd_workspace_bytes(int M, int C);
int dfw_layernorm_bwd(const void* x, const void* dy, int dtype, const float* gamma, void* dx, float* dgamma, float* dbeta,
                      int M, int C, float eps, void* workspace, void* stream);
int dfw_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_f16, int M, int C,
                  float eps, void* stream);

/* Row softmax for the VAE mid-block attention (single head, d=512): s fp32 [M, L] -> p bf16 (fp16 if y_f16) [M, L],
 * p = softmax(s * scale).  ref: AutoencoderKL mid_block Attention (upstream), pipeline:852,901. */
int dfw_softmax_rows(const float* s, void* p, int y_f16, int M, int L, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 evaluation mode (the reference's shipped eval numerics: evaluation_util/main_oss.py:332-336 is fp32-only).
 * GEMMs / convolutions run on the entry points above with SPLIT operands: x = hi + lo (two 16-bit values),
 * x.w ~= hi_x hi_w + lo_x hi_w + hi_x lo_w = one GEMM over a 3x longer channel axis.  What surrounds them in fp32:
 *   dfw_split3_16      x fp32 [rows, C] (row stride in floats) -> y 16-bit [rows, 3C]; role 0 = [hi | lo | hi] (activation),
 *                      role 1 = [hi | hi | lo] (weight side of an activation x activation product); bf16, or fp16 if y_f16
 *   dfw_groupnorm_f32  GroupNorm (+ SiLU, exact sigmoid) x, y fp32 [N, HW, C], two-pass statistics
 *   dfw_layernorm_f32  x, y fp32 [M, C]
 *   dfw_softmax_rows_f32  p = softmax(s * scale) per row, fp32 [M, L] (in place allowed)
 *   dfw_geglu_f32      h fp32 [rows, 2F] = (value | gate) -> y fp32 [rows, F] = value * gelu_erf(gate)
 *   dfw_attn_f32       softmax(q [k_self ; k_bank]^T scale) [v_self ; v_bank], head dim 64, fp32 on CUDA cores; strides in
 *                      floats, k / v pointers 16-byte aligned with strides % 4 == 0; a batch stride of 0 shares K / V
 *                      (cross-attention to one prompt).  ref: attention_processor.py:251-271. */
int dfw_split3_16(const float* x, long long x_row_stride, void* y, long long rows, int C, int role, int y_f16, void* stream);
int dfw_groupnorm_f32(const float* x, const float* gamma, const float* beta, float* y, int N, int HW, int C, int groups,
                      float eps, int apply_silu, void* stream);
int dfw_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, long long M, int C, float eps,
                      void* stream);
int dfw_softmax_rows_f32(const float* s, float* p, int M, int L, float scale, void* stream);
int dfw_geglu_f32(const float* h, float* y, long long rows, int F, void* stream);
int dfw_attn_f32(const float* q, long long q_batch_stride, long long q_row_stride, const float* k_self, const float* v_self,
                 long long kv_batch_stride, long long kv_row_stride, const float* k_bank, const float* v_bank,
                 long long bank_batch_stride, long long bank_row_stride, float* o, long long o_batch_stride,
                 long long o_row_stride, int B, int heads, int Lq, int Ls, int Lb, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Layout / small-channel helpers (CUDA cores, bandwidth-bound).
 * ------------------------------------------------------------------------------------------------------------ */
/* nearest 2x upsample NHWC [N,H,W,C] (16-bit copied as is, or fp32 if x_f32 -> bf16 / fp16 per y_f16)
 * -> 16-bit [N,2H,2W,C].  C % 8 == 0.
 * ref: Upsample2D F.interpolate(scale 2, nearest) (upstream) */
int dfw_upsample2x_nhwc(const void* x, int x_f32, void* y, int y_f16, int N, int H, int W, int C, void* stream);
/* channel concat: y[rows,Ca+Cb] = cat(a[rows,Ca], b[rows,Cb]), elem_bytes 2 (bf16) or 4 (fp32).
 * ref: torch.cat([hidden_states, res_hidden_states], dim=1) in the UNet up blocks (upstream, unet:1226) */
int dfw_concat_channels(const void* a, const void* b, void* y, long long rows, int Ca, int Cb, int elem_bytes,
                        void* stream);
/* fp32 -> bf16 / fp16 cast (n % 8 == 0): the fp32 residual stream becomes an MMA operand (shortcut / downsample). */
int dfw_cast_f32_to_16(const float* x, void* y, int y_f16, long long n, void* stream);
/* 3x3 / stride 1 / pad 1 convolution with a tiny input channel count (Cin <= 8), CUDA cores:
 *   x  fp32 NCHW [N,Cin,H,W] (the reference's image / latent layout);  w fp32 [Cout,3,3,Cin]; bias fp32 [Cout]
 *   y  NHWC [N,H,W,Cout] with y_dtype 0 = bf16, 1 = fp32, 2 = fp16; Cout % 64 == 0, Cout <= 512.
 * ref: UNet conv_in / conv_in_ref (unet_2d_condition.py:301-306,1118-1121), VAE encoder conv_in (3->128),
 *      VAE decoder conv_in (4->512) (upstream). */
int dfw_conv3x3_small_cin(const float* x, const float* w, const float* bias, void* y, int y_dtype, int N, int H,
                          int W, int Cin, int Cout, void* stream);
/* im2col for the same tiny-Cin 3x3 convolutions (Cin <= 16): x fp32 NCHW -> y 16-bit rows [N*H*W, Kpad],
 * k = (kh*3+kw)*Cin + c, zero padded to Kpad (% 64 == 0); the convolution then is dfw_linear with K = Kpad and the
 * weight matrix [Cout, Kpad] in the same k order (tensor cores instead of CUDA cores for the 3 -> 128 @ 512^2 layer). */
int dfw_im2col3x3_small(const float* x, void* y, int y_f16, int N, int H, int W, int Cin, int Kpad, void* stream);
/* 1x1 conv on <= 8 channels, fp32 math, arbitrary element strides (so it also converts NHWC <-> NCHW):
 *   y[n,p,co] = (sum_ci w[co,ci] * (x[n,p,ci] * in_scale) + b[co]) * out_scale
 *   x element (n,p,ci) at x[n*x_ns + p*x_ps + ci*x_cs]; y likewise.  w [Cout,Cin] and b [Cout] are HOST pointers
 *   (<= 72 floats, passed to the kernel by value).
 * ref: vae.quant_conv / post_quant_conv and the latent scale factors (pipeline:852-861, 898-901) and the
 *      DDIM step collapse z0 = -eps (marigold/util/scheduler_customized.py:107-180). */
int dfw_pointwise_small(const float* x, long long x_ns, long long x_ps, long long x_cs, const float* w_host,
                        const float* b_host, float in_scale, float out_scale, float* y, long long y_ns,
                        long long y_ps, long long y_cs, int N, int HW, int Cin, int Cout, void* stream);
/* fp32 NHWC rows [N*HW, x_row_stride] (first C used) -> fp32 NCHW [N,C,HW]: y = clamp(x*scale + shift, lo, hi). */
int dfw_nhwc_f32_to_nchw_f32(const float* x, int x_row_stride, float* y, int N, int C, int HW, float scale,
                             float shift, float lo, float hi, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Segmentation tail.
 * ------------------------------------------------------------------------------------------------------------ */
/* seg post-processing: dec fp32 NHWC rows [N*HW, row_stride] (3 used) ->
 *   seg_f32 NCHW [N,3,H,W] = (clip(x,-1,1)*0.5+0.5)*255         (single_infer return value, pipeline:787-795)
 *   seg_u8  NCHW [N,3,H,W] = uint8 truncation of clip(seg,0,255) (pipeline:534)            (either may be NULL) */
int dfw_seg_post(const float* dec, int row_stride, float* seg_f32, uint8_t* seg_u8, int N, int HW, void* stream);

/* K10: fused VAE decoder head -- GroupNorm-apply + SiLU + conv3x3 128->3 + clip(-1,1) + (.)*0.5+0.5 + (.)*255 + uint8
 * truncation in ONE pass over x (ref: diffews/marigold_pipeline_rgb_latent_noise.py:887-905 decode_seg -> upstream
 * Decoder tail conv_norm_out / SiLU / conv_out, :787-795, :534).  x: 16-bit [N,H,W,128]; scale_shift: fp32 [N,2,128]
 * from dfw_gn_scale_shift; wb: device copy of dfw_seg_head_prepare_weights' output; bias_host: 3 floats (host);
 * out_u8 [N,3,H,W] and/or out_f32 [N,3,H,W] (float in [0,255]).  Requires W % 16 == 0. */
long long dfw_seg_head_weight_u32(void);
int dfw_seg_head_prepare_weights(const float* w_oihw_host, int f16, uint32_t* out_host);
int dfw_seg_head_u8(const void* x, int f16, const float* scale_shift, const uint32_t* wb, const float* bias_host,
                    uint8_t* out_u8, float* out_f32, int N, int H, int W, void* stream);

/* K8  reverse-threshold binarisation + intersection/union histogram, one launch, integer counts.
 * ref: evaluation_util/main_oss.py:128-134 (to_tensor, max()*r, mean(dim=1) > thr — CPU fp32 semantics reproduced
 *      bit-exactly: ((R/255 + G/255) + B/255) / 3  >  fp32(max/255) * r), evaluation_util/common/evaluation.py:12-39
 *      (histc bins=2 on pred[pred==gt], pred, gt; ignore_index 255), logger.py:35-37 (accumulation, here int64).
 *   pred_u8 [B,3,H,W] uint8 (or, if pred_is_mask, an already binarised [B,H,W] {0,1} mask: the thresholding is
 *   skipped and the call is exactly Evaluator.classify_prediction); gt [B,H,W] uint8 in {0,1};
 *   ignore [B,H,W] uint8 {0,1} or NULL;
 *   r_threshold (0.25); the max is taken PER EPISODE (the reference only runs bsz=1, where they coincide).
 *   out: area_inter int64 [B,2], area_union int64 [B,2] (bin 0 background, bin 1 foreground);
 *        mask_out uint8 [B,H,W] (0/1, 255 where ignored) or NULL.
 *   workspace: dfw_rthres_workspace_bytes(B) bytes (zeroed by the call). */
long long dfw_rthres_workspace_bytes(int B);
int dfw_rthres_iou_hist(const uint8_t* pred_u8, int pred_is_mask, const uint8_t* gt, const uint8_t* ignore,
                        float r_threshold, long long* area_inter, long long* area_union, uint8_t* mask_out, int B,
                        int H, int W, void* workspace, void* stream);
/* AverageMeter.update: buf[2, nclass] (int64) += counts[b, :] at column class_id[b].  ref: logger.py:35-37 */
int dfw_iou_accumulate(const long long* area_inter, const long long* area_union, const long long* class_id,
                       long long* inter_buf, long long* union_buf, int B, int nclass, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K9  Episode preprocessing (the data format in front of the hot path).
 * ref: evaluation_util/data/dataset.py:36-40 (Resize((S,S)) -> ToTensor -> Normalize([0.5],[0.5])),
 *      evaluation_util/data/coco.py:38-47, :92-93, pascal.py:42-55, :78-83, fss.py:38-47, :80-84.
 * All decoded images of a batch sit in ONE device buffer `base`; `descs` (device, n entries, 8-byte aligned, may live
 * inside the same buffer) says where.  Results are byte-identical to the reference's CPU path: Pillow's two-pass
 * fixed-point bilinear resample (ImagingResample, uint8 intermediate) and torch's nearest index rule.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct DfwImageDesc {
    long long offset; /* byte offset of pixel (0,0) from `base`                                        */
    int h, w;         /* source height / width                                                         */
    int row_stride;   /* bytes per source row (>= 3*w for RGB images, >= w for label masks)            */
    int param;        /* dfw_mask_nearest mode 0: the label value that maps to 1 (class + 1)           */
} DfwImageDesc;

/* bytes of scratch dfw_resize_normalize_u8 needs for n images no larger than max_h x max_w (-1 on bad arguments) */
long long dfw_preproc_workspace_bytes(int n, int max_h, int max_w, int out_h, int out_w);
/* RGB uint8 HWC images -> dst_f32 [n,3,out_h,out_w] = ((resized / 255) - mean) / std  (fp32, the reference's op order)
 * and / or dst_u8 [n,out_h,out_w,3] = the resized bytes (== PIL.Image.resize((out_w,out_h), BILINEAR)); either may be
 * NULL.  3 launches for the whole batch (coefficients, horizontal pass, vertical pass + normalise). */
int dfw_resize_normalize_u8(const void* base, const void* descs, int n, int max_h, int max_w, float* dst_f32,
                            uint8_t* dst_u8, int out_h, int out_w, float mean, float std, void* workspace,
                            long long workspace_bytes, void* stream);
/* uint8 label masks [h,w] -> mask_out fp32 [n,out_h,out_w] in {0,1} at F.interpolate(mode="nearest") positions;
 * mode 0: label == desc.param; mode 1: label >= 128.  boundary_out (NULL ok) = floor(label / 255) (PASCAL ignore). */
int dfw_mask_nearest(const void* base, const void* descs, int n, float* mask_out, float* boundary_out, int out_h,
                     int out_w, int mode, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K10  Training-step tail (first pieces): MSE loss, global gradient-norm clipping, fused AdamW.
 * ref: train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1384 (F.mse_loss(pred.float(), target.float())),
 *      :1393 (clip_grad_norm_), :1186-1194 + :1394 (torch.optim.AdamW(...).step()).
 * One launch covers every parameter tensor: `tensors` is a DEVICE array of DfwAdamTensor; the work is cut into chunks of
 * `chunk_elems` (% 4 == 0) elements, chunk c = elements [chunk_offset[c], +chunk_elems) of tensor chunk_tensor[c]
 * (both DEVICE arrays of n_chunks entries).  fp32 parameters / gradients / moments.  Deterministic (no fp atomics).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct DfwAdamTensor {
    void* param;        /* fp32 [numel], updated in place                                    */
    const void* grad;   /* fp32 [numel]                                                      */
    void* exp_avg;      /* fp32 [numel]                                                      */
    void* exp_avg_sq;   /* fp32 [numel]                                                      */
    void* param16;      /* NULL or 16-bit [numel]: receives the rounded updated parameter    */
    long long numel;
} DfwAdamTensor;

/* norm_out[0] = sqrt(sum of all grad^2); coef_out[0] = min(1, max_norm / (norm + 1e-6)) (torch clip_grad_norm_, L2);
 * partial: n_chunks floats of scratch.  The gradients are NOT modified: pass coef_out as dfw_adamw_step's grad_scale. */
int dfw_grad_norm_clip_coef(const void* tensors, const int* chunk_tensor, const long long* chunk_offset, int n_chunks,
                            int chunk_elems, float max_norm, float* partial, float* norm_out, float* coef_out,
                            void* stream);
/* torch.optim.AdamW single-tensor arithmetic: p *= 1 - lr*wd; m += (1-b1)(g-m); v = v*b2 + (1-b2) g*g;
 * p += -(lr/(1-b1^step)) * m / (sqrt(v)/sqrt(1-b2^step) + eps), with g = grad * grad_scale[0] (device scalar, NULL = 1).
 * p16_format 0: no 16-bit copy; 1: bf16; 2: fp16 into DfwAdamTensor.param16.  step counts from 1.  The hyper-parameters
 * are doubles because torch derives 1-beta, 1-lr*wd and the bias corrections from Python floats before rounding to fp32. */
int dfw_adamw_step(const void* tensors, const int* chunk_tensor, const long long* chunk_offset, int n_chunks,
                   int chunk_elems, double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                   const float* grad_scale, int p16_format, void* stream);
/* loss_out[0] = mean((pred - target)^2); dpred (NULL ok) = upstream * 2 (pred - target) / n.
 * workspace: dfw_mse_workspace_floats() floats. */
/* GEGLU backward (diffusers GEGLU: v, g = proj(x).chunk(2, -1); y = v * gelu_erf(g); FeedForward of BasicTransformerBlock,
 * upstream): h [M, 2F] = (value | gate) pre-activations, dy [M, F] -> dh [M, 2F]; one dtype (0 bf16 / 1 fp32 / 2 fp16),
 * F % 8 == 0. */
int dfw_geglu_bwd(const void* h, const void* dy, void* dh, int dtype, long long M, int F, void* stream);
long long dfw_mse_workspace_floats(void);
int dfw_mse_loss(const float* pred, const float* target, long long n, float upstream, float* loss_out, float* dpred,
                 float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Gradient kernels of the training step (csrc/grad.cu).
 * ref: train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1386 accelerator.backward(loss): torch autograd through
 *      every nn.Conv2d / nn.Linear / Upsample2D / torch.cat / GEGLU of the UNet (cuDNN wgrad + dgrad, cuBLAS) (upstream).
 *
 * Weight gradient of y = conv(x, w) (ksize 1|3, stride 1|2 with pad (ksize-1)/2; a Linear is N = Hin = 1, Win = M):
 *   dw[co, (kh*ks+kw)*Cin + ci] (=|+=) scale * sum over output pixels of dy[p, co] * x[p*stride + (kh,kw) - pad, ci]
 *   x 16-bit [N,Hin,Win,Cin] (Cin % 32 == 0), dy 16-bit [N,Hin/stride,Win/stride,Cout] (Cout % 8 == 0), dw fp32
 *   [cout_store, ks*ks*Cin] (cout_store <= Cout: rows beyond it are not written — dy channel-padded for tiny Cout).
 * tcgen05 GEMM with the pixel index as the contraction: both operands are read in place (MN-major TMA tiles), split over
 * pixel ranges across CTAs, partials summed in fixed order (deterministic).  workspace: dfw_conv_wgrad_workspace_bytes. */
long long dfw_conv_wgrad_workspace_bytes(int N, int Hin, int Win, int Cin, int Cout, int ksize, int stride);
int dfw_conv_wgrad(const void* x, const void* dy, float* dw, int N, int Hin, int Win, int Cin, int Cout, int cout_store,
                   int ksize, int stride, int f16, float scale, int accumulate, void* workspace, void* stream);
/* out[c, t', r] = tap_map[t'] >= 0 ? w[r, tap_map[t'], c] : 0;  w 16-bit [R, T, C] -> out 16-bit [C, T_out, R], T_out <= 16.
 * Builds the operand of the data-gradient convolution from a forward weight: Linear transpose (T = 1), 3x3 rotation
 * (tap_map[t] = 8 - t), the four phase filters of the stride-2 transposed convolution. */
int dfw_weight_permute(const void* w, void* out, int R, int T, int C, int T_out, const int* tap_map, void* stream);
/* out[g, c] (=|+=) scale * sum over the rows of group g of x[g*rows_per_group + r, c]: bias gradients (groups = 1) and
 * per-image time-embedding gradients (groups = N).  dtype 0 bf16 / 1 fp32 / 2 fp16; C % 8 == 0; deterministic.
 * workspace: groups * dfw_colsum_chunks(rows_per_group, groups) * C floats. */
int dfw_colsum_chunks(long long rows_per_group, int groups);
int dfw_colsum(const void* x, int dtype, float* out, long long rows_per_group, int groups, int C, float scale, int accumulate,
               float* workspace, void* stream);
/* backward of the nearest-2x upsample: dx[n,h,w,:] = sum of the 2x2 block of dy [N,2H,2W,C]; 16-bit, C % 8 == 0. */
int dfw_downsum2x_nhwc(const void* dy, void* dx, int f16, int N, int H, int W, int C, void* stream);
/* backward of the channel concat: a = y[:, :Ca], b = y[:, Ca:]; 16-bit rows, Ca % 8 == Cb % 8 == 0. */
int dfw_split_channels(const void* y, void* a, void* b, long long rows, int Ca, int Cb, void* stream);
/* y[M, F] = h[:, :F] * gelu_erf(h[:, F:]) on 16-bit (value | gate) pre-activations kept for dfw_geglu_bwd. */
int dfw_geglu_fwd(const void* h, void* y, int f16, long long M, int F, void* stream);
/* y[n,h,w,c] = c < C ? scale * x[n,c,h,w] : 0: fp32 NCHW -> 16-bit NHWC with Cpad channels (d loss / d prediction). */
int dfw_nchw_f32_to_nhwc16_pad(const float* x, void* y, int N, int C, int H, int W, int Cpad, float scale, int f16,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFEWS_B200_H_ */
