/*
 * magicdance_b200 — C ABI of the B200 (sm_100a) kernels behind MagicPose's DDIM denoising hot path.
 *
 * The reference (Boese0601/MagicDance) has no FFI: its "plugin API" is Python classes looked up by
 * YAML `target:` strings (model_lib/ControlNet/ldm/util.py:72-87).  The drop-in Python classes in
 * this repo (model_lib/ControlNet/cldm/cldm.py, .../ldm/...) keep those names and signatures and
 * call THIS library for every per-step tensor op.  Each entry point below names the reference
 * interface (file:line, relative to /root/reference/model_lib/ControlNet/) it replaces.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers (cudaMalloc'ed / torch CUDA storage) unless noted;
 *   - activations are fp16, channels-last: an NCHW tensor (B,C,H,W) is stored as [B][H][W][C],
 *     which is also the token-major (B, H*W, C) matrix the transformer blocks use;
 *   - weights are fp16, "K-major": Linear (out,in) as is; Conv2d OIHW repacked to [O][kh][kw][I];
 *   - `stream` is a cudaStream_t (0 = legacy default stream); every call is asynchronous;
 *   - return value: 0 on success, negative MDB_ERR_* otherwise; mdb_last_error() gives the text.
 *   - there is NO CPU fallback: without an sm_100 device every compute entry returns an error.
 */
#ifndef MAGICDANCE_B200_H_
#define MAGICDANCE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_OK 0
#define MDB_ERR_INVALID (-1)
#define MDB_ERR_CUDA (-2)
#define MDB_ERR_UNSUPPORTED (-3)

#define MDB_ABI_VERSION 2

typedef void* mdb_stream_t;

/* library plumbing */
int mdb_abi_version(void);
const char* mdb_last_error(void);
/* 0 if the current CUDA device is compute capability 10.x, MDB_ERR_UNSUPPORTED otherwise */
int mdb_device_check(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t mdb_launch_count(void);
/* sizeof(mdb_gemm_desc) (which = 0) / sizeof(mdb_attn_desc) (which = 1): a binding checks its struct mirrors */
int64_t mdb_abi_struct_bytes(int32_t which);

/* Launch heuristics, process-wide.  The defaults are what the B200 measurements selected (profiles/); tests use the
 * setter to force a kernel variant onto small problems. */
#define MDB_TUNE_GEMM_PAIR_MIN_TILES 1 /* grids of >= this many 128-row tiles use the persistent CTA-pair GEMM (128) */
#define MDB_TUNE_ATTN40_2Q_MIN_CTAS 3  /* d=40 attention grids of >= this many CTAs use the two-Q-tile kernel (2048) */
#define MDB_TUNE_GEMM_BN80_BELOW 4     /* N %% 160 == 0 layers with fewer 160-wide CTAs than this take 80-wide tiles (100) */
int mdb_set_tuning(int32_t key, int32_t value);
int32_t mdb_get_tuning(int32_t key);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM / implicit-GEMM convolution (tcgen05 + TMEM + TMA).
 *   D[M,N] = epilogue( A[M,K] * B[N,K]^T )     fp16 in, fp32 accumulate, fp16 out
 * Replaces: nn.Linear in CrossAttention.to_q/to_k/to_v/to_out (ldm/modules/attention.py:154-161),
 * GEGLU.proj / FeedForward.net[2] (attention.py:53-56,68-72), the 1x1 convs proj_in/proj_out
 * (attention.py:342-361), ResBlock.skip_connection and the ControlNet zero convs
 * (openaimodel.py:254-261, cldm.py:733-734), and — in conv mode — every 3x3 stride-1 pad-1
 * conv_nd of ResBlock / Upsample / Downsample-after-im2col (openaimodel.py:225,249-252,127,175).
 * ---------------------------------------------------------------------------------------------- */
#define MDB_EPI_NONE 0
#define MDB_EPI_GEGLU 1 /* B/bias rows interleaved [32 value | 32 gate]; D is [M][N/2] = v*gelu_erf(g) */

typedef struct mdb_gemm_desc {
  const void* a;   /* plain: fp16 [M][k1], row stride lda.  conv: NHWC fp16 [nb][h][w][c]        */
  int64_t lda;
  const void* a2;  /* optional 2nd source for K columns [k1, K) (fused torch.cat along channels) */
  int64_t lda2;
  int32_t k1;      /* columns taken from a; == k when a2 == NULL                                  */
  int32_t conv;    /* 0 plain; 1 | 2 = 3x3 pad-1 implicit GEMM with stride 1 | 2 over the NHWC input
                      (K = 9*c, M = nb*ho*wo, ho = (h-1)/stride+1): no im2col buffer (TMA element strides) */
  int32_t nb, h, w, c;
  const void* b;   /* fp16 [N][K], row stride ldb                                                 */
  int64_t ldb;
  void* d;         /* fp16 [M][N] (GEGLU: [M][N/2]), row stride ldd                               */
  int64_t ldd;
  const float* bias;          /* fp32, may be NULL; bias[(row / rows_per_batch) * bias_batch_stride + col] */
  int64_t bias_batch_stride;  /* 0 => one bias row for all batches                                */
  int32_t rows_per_batch;     /* rows of D per batch element (ignored when bias_batch_stride == 0) */
  int32_t epilogue;           /* MDB_EPI_*                                                        */
  const void* residual;       /* fp16 [M][N] added after bias, may be NULL                        */
  int64_t ldr;
  int32_t m, n, k;
  int32_t splits;             /* 0: automatic (1, 2, 4 or 8, in-cluster reduction); 1: none; >1: as given */
  float* splitk_ws;           /* fp32 [splits][M][N] scratch, only for explicit split counts other than 2, 4, 8 */
  /* LayerNorm over A's rows folded into the GEMM (norm2 -> attn2.to_q of BasicTransformerBlock,
   * attention.py:271,312-314): with B = W diag(gamma) and bias = W beta (+ the layer's own bias) supplied by the
   * caller, ln_u[n] = sum_k B[n][k] makes D = rstd_r (A B^T - mean_r ln_u) + bias equal to LayerNorm(A) W^T + b; K must
   * be the normalised width.  Small grids only: every N tile recomputes the statistics of its rows (no split-K, no
   * GEGLU epilogue, single-CTA tiles).  NULL = off. */
  const float* ln_u;
  float ln_eps;
} mdb_gemm_desc;

int mdb_gemm_f16(const mdb_gemm_desc* desc, mdb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention, FlashAttention-style tile loop on tcgen05, with TWO key/value sources whose
 * keys are concatenated in-kernel:  out = softmax([Q K0^T | Q K1^T] * scale) [V0 ; V1].
 * Replaces CrossAttention._forward (attention.py:168-199) / MemoryEfficientCrossAttention
 * (attention.py:225-250) AND the torch.cat([x_norm1] + bank) of BasicTransformerBlock 'read'
 * mode (attention.py:303-307): source 0 = the layer's own tokens, source 1 = the appearance bank.
 *   q   : fp16 [B*Nq][heads*d]                    (row stride ldq)
 *   k0  : fp16 [kv0_batches*N0][heads*d]          (row stride ldk0); kv0_batches is B or 1 (shared)
 *   vt0 : fp16 [heads*d][kv0_batches*ldv0_batch]  V TRANSPOSED: row = channel, col = key; each
 *         batch occupies ldv0_batch columns (>= N0, multiple of 8), row stride ldvt0
 *   k1/vt1 : same for source 1 (NULL / N1 = 0 when absent); only batches b < bank_batches use it
 *   out : fp16 [B*Nq][heads*d] (row stride ldo)
 * d in {40, 80, 160}.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mdb_attn_desc {
  const void* q; int64_t ldq;
  const void* k0; int64_t ldk0; const void* vt0; int64_t ldvt0; int32_t n0; int32_t kv0_batches; int32_t ldv0_batch;
  const void* k1; int64_t ldk1; const void* vt1; int64_t ldvt1; int32_t n1; int32_t kv1_batches; int32_t ldv1_batch;
  void* out; int64_t ldo;
  int32_t batch, heads, d, nq;
  int32_t bank_batches;
  float scale; /* d^-0.5 (attention.py:152) */
} mdb_attn_desc;

int mdb_attention_f16(const mdb_attn_desc* desc, mdb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm(32 groups, affine) [+ SiLU], fp32 statistics, over channels-last fp16; optionally the
 * input is the channel-concatenation of two tensors (fuses torch.cat([h, hs.pop()], 1),
 * cldm.py:104).  Replaces GroupNorm32 (ldm/modules/diffusionmodules/util.py:252-254) + nn.SiLU in
 * ResBlock.in_layers/out_layers and UNet.out (openaimodel.py:222-226,246-248,744-748), and
 * Normalize (attention.py:89-90, eps 1e-6) in SpatialTransformer.
 *   x1 [B][hw][c1], x2 [B][hw][c2] (x2 NULL => c2 = 0), y [B][hw][c1+c2].
 * Deterministic (fixed-order reductions, no atomics on data) and pivot-shifted (sums of x - x[b,0,first channel of
 * the group]), so large-mean activations do not cancel.  Two paths, chosen by batch size (mode 0), or forced
 * (mode 1 = two kernels, mode 2 = cluster):
 *   - ONE launch: a thread-block cluster per (batch element, group) exchanges its partial sums through
 *     distributed shared memory (channels per group even, i.e. c a multiple of 64); ws may be NULL;
 *   - stats (+ last-CTA fold) -> apply: needs ws of mdb_groupnorm_ws_floats(c, batch, hw) floats that were ZERO
 *     when first used (self-resetting tickets; calls of any shape on one stream may share one workspace).
 * ---------------------------------------------------------------------------------------------- */
int mdb_groupnorm_f16(const void* x1, int32_t c1, const void* x2, int32_t c2, const float* gamma, const float* beta,
                      void* y, float* ws, int32_t batch, int32_t hw, float eps, int32_t silu, int32_t mode,
                      mdb_stream_t stream);
int64_t mdb_groupnorm_ws_floats(int32_t c, int32_t batch, int32_t hw);

/* LayerNorm over the last dim (eps 1e-5), fp16 [rows][c] -> fp16; replaces nn.LayerNorm norm1/2/3 of
 * BasicTransformerBlock (attention.py:270-272). */
int mdb_layernorm_f16(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int32_t c,
                      float eps, mdb_stream_t stream);

/* Generic direct 3x3 conv (pad 1, stride 1|2) for shapes the tensor-core path does not take
 * (cin or cout not a multiple of 64): the ControlNet hint encoder (cldm.py:599-615), the 4->320
 * input conv (openaimodel.py:554-558) and the 320->4 output conv (openaimodel.py:744-748).
 *   x NHWC fp16 [B][h][w][cin]; wt fp16 [cout][3][3][cin]; y NHWC fp16 [B][ho][wo][cout];
 *   y = act(conv(x) + bias) + residual, act = SiLU when silu != 0, residual (same shape as y) optional
 *   (the ControlNet's `h += guided_hint`, cldm.py:745-749). */
int mdb_conv3x3_direct_f16(const void* x, const void* wt, const float* bias, const void* residual, void* y,
                           int32_t batch, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t stride,
                           int32_t silu, mdb_stream_t stream);

/* im2col for 3x3 pad-1 convolutions, stride 1 or 2: x NHWC [B][h][w][c] -> col [B*ho*wo][9*c] with K order
 * (kh, kw, c), ho = (h-1)/stride+1, consumed by mdb_gemm_f16.  Used for Downsample.op (stride 2,
 * openaimodel.py:154-180) and as the general path for latent sizes whose rows do not tile into the
 * 128-pixel TMA boxes of the implicit-GEMM conv (any image size the reference accepts works). */
int mdb_im2col3x3_f16(const void* x, void* col, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t stride,
                      mdb_stream_t stream);

/* The same gather for a window anchored at the output pixel with ONE padding row / column at the bottom / right
 * only: ho = (h + 1 - 3)/stride + 1.  Replaces F.pad(x, (0,1,0,1)) + Conv2d(k=3, stride=2, padding=0) of the
 * first-stage VAE encoder's Downsample (ldm/modules/diffusionmodules/model.py:71-90) together with mdb_gemm_f16. */
int mdb_im2col3x3_br_f16(const void* x, void* col, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t stride,
                         mdb_stream_t stream);

/* nearest x2 upsample (Upsample.forward, openaimodel.py:129-139): NHWC [B][h][w][c] -> [B][2h][2w][c] */
int mdb_upsample2x_f16(const void* x, void* y, int32_t batch, int32_t h, int32_t w, int32_t c, mdb_stream_t stream);

/* y = a + b (b broadcast over the batch when b_batches == 1); the ControlNet residual adds
 * `h += pose_control.pop()` / `hs.pop() + pose_control.pop()` (cldm.py:93-104). n = elements per batch */
int mdb_add_f16(const void* a, const void* b, void* y, int64_t n_per_batch, int32_t batch, int32_t b_batches,
                mdb_stream_t stream);

/* timestep_embedding (util.py:189-209): t int64 [t_count] -> fp32 [batch][dim], [cos | sin], max_period 1e4; row b
 * uses t[b % t_count] (t_count = batch: one timestep per sample; fewer: the timesteps repeat, e.g. cond | uncond) */
int mdb_timestep_embedding_f32(const int64_t* t, int32_t t_count, float* out, int32_t batch, int32_t dim,
                               mdb_stream_t stream);

/* Skinny Linear for the timestep path (rows <= 16): out[r][n] = sum_k f(x[r][k]) W[n][k] + bias[n],
 * f = SiLU when silu_in, fp32 in/out, fp16 weights.  Replaces time_embed (openaimodel.py:547-551)
 * and every ResBlock.emb_layers (openaimodel.py:238-244) — all 22 of a network in one launch by
 * stacking their weights along n.  silu_out applies SiLU to the result (time_embed's middle SiLU). */
int mdb_skinny_linear_f32(const float* x, const void* w, const float* bias, float* out, int32_t rows, int32_t n,
                          int32_t k, int32_t silu_in, int32_t silu_out, mdb_stream_t stream);


/* Row softmax in place over fp16 logits x[rows][cols] (row pitch ld elements), fp32 arithmetic:
 * x <- softmax(scale * x) along the columns.  The single-head, 512-channel attention of the first-stage VAE's
 * middle block (ldm/modules/diffusionmodules/model.py:186-193: torch.bmm -> * c^-0.5 -> softmax) is run as
 * GEMM -> this -> GEMM, its head dimension being outside the fused attention kernel's range (<= 160). */
int mdb_softmax_rows_f16(void* x, int64_t ld, int32_t rows, int32_t cols, float scale, mdb_stream_t stream);

/* layout/precision boundary: the reference passes NCHW fp32 tensors (cldm.py:1099) */
/* y holds the batch `copies` times over ([copies*batch][h][w][c]): the paired cond | uncond batch of p_sample_ddim */
int mdb_nchw_f32_to_nhwc_f16(const float* x, void* y, int32_t batch, int32_t c, int32_t h, int32_t w, int32_t copies,
                             mdb_stream_t stream);
int mdb_nhwc_f16_to_nchw_f32(const void* x, float* y, int32_t batch, int32_t c, int32_t h, int32_t w, mdb_stream_t stream);

/* CFG combine + DDIM update in one pass (ddim.py:605,617-645; eps-parameterisation):
 *   e = e_u + scale (e_c - e_u); pred_x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);
 *   x_prev = sqrt(a_prev) pred_x0 + sqrt(1 - a_prev - sigma^2) e + sigma * noise   (noise may be NULL when sigma == 0)
 * all tensors fp32, n elements.  coef is a DEVICE array of 6 floats
 *   {scale, sqrt(a_t), sqrt(a_prev), sqrt(1 - a_prev - sigma^2), sigma, sqrt(1 - a_t)}
 * so that one captured CUDA graph serves every DDIM step.  update_x != 0: x is overwritten with x_prev as well
 * (the chain's state advances in place). */
int mdb_cfg_ddim_update_f32(float* x, const float* eps_c, const float* eps_u, const float* noise, float* x_prev,
                            float* pred_x0, int64_t n, const float* coef, int32_t update_x, mdb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGICDANCE_B200_H_ */
