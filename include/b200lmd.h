/* b200lmd — C ABI of the B200-native layout-grounded denoising path.
 *
 * The reference (TonyLianLong/LLM-groundedDiffusion) is pure Python and has no FFI of its own; these entry points are
 * what a binding for its hot path attaches to.  Each block cites the reference interface it replaces.
 *
 * Conventions: every function returns 0 on success, non-zero on failure (b200lmd_last_error() gives the text); no
 * exceptions cross the boundary; all pointers are DEVICE pointers unless a name ends in _host; the caller owns every
 * buffer it passes (the library owns only its weight copies and workspace arena); all work is asynchronous on the
 * cudaStream_t passed as `void* stream` (0 = default stream); handles are not thread-safe.
 * Activations are fp16, token/pixel-major ("NHWC": [B, H, W, C] == [B*n, C]); accumulation is fp32.
 */
#ifndef B200LMD_H
#define B200LMD_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* b200lmd_last_error(void);
int b200lmd_version(void);
/* runtime switches for A/B measurements of kernel generations ("attn_v2": 1 = ping-pong online-softmax attention) */
int b200lmd_set_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------------ dense ops
 * y[M,N] = alpha * x[M,K] . W[N,K]^T (+ bias[N]) (+ residual[M,N]).  Replaces nn.Linear / 1x1 nn.Conv2d call sites:
 * models/attention_processor.py:127-142 (to_q/k/v/out), models/transformer_2d.py:146-150,205-207 (proj_in/out),
 * models/attention.py:281,325 (FeedForward).  x,W,residual,y fp16; bias fp32 (may be NULL). ld* in elements. */
int b200lmd_linear_f16(const void* x, int ldx, const void* w, const void* bias, const void* residual, int ldr,
                       void* y, int ldy, void* y_f32, int M, int N, int K, float alpha, int accumulate, void* stream);

/* GEGLU: y[M,F] = (xW_v^T + b_v) * gelu_erf(xW_g^T + b_g), W = [2F, K] (value rows then gate rows, as the
 * reference's GEGLU.proj stores them, models/attention.py:323-335).  `w_il`/`bias_il` must be the tile-interleaved
 * copies made by b200lmd_geglu_interleave_*.  pre (optional) receives the pre-activation [M, 2F] (interleaved). */
int b200lmd_geglu_interleave_w(const void* w, void* w_il, int F, int K, void* stream);
int b200lmd_geglu_interleave_b(const void* bias, void* bias_il, int F, void* stream);
int b200lmd_linear_geglu_f16(const void* x, int ldx, const void* w_il, const void* bias_il, void* y, void* pre, int M,
                             int F, int K, void* stream);

/* 3x3 / stride 1 / pad 1 convolution on NHWC fp16: y[B,H,W,Cout] = conv(x[B,H,W,Cin], w[Cout,3,3,Cin]) + bias
 * (+ chan_add[B,Cout], the time-embedding projection) (+ residual).  Replaces the conv1/conv2 of diffusers-0.18
 * ResnetBlock2D reached from models/unet_2d_blocks.py:23 and conv_in/conv_out at models/unet_2d_condition.py:289,567.
 * w is [Cout][9][Cin] fp16 (tap = ky*3+kx); bias/chan_add fp32 or NULL. */
int b200lmd_conv3x3_f16(const void* x, const void* w, const void* bias, const void* chan_add, const void* residual,
                        void* y, void* y_f32, int B, int H, int W, int Cin, int Cout, void* stream);

/* Descriptor form of the implicit GEMM behind every op above:
 *   D[m, n] = sum_{tap, c} A[pixel(m) + off(tap), c] * W[n, wtap(tap), c]
 * A is NHWC fp16 [aB, aH, aW, a_ld] (channels [0, Cin) are contracted); the output-pixel grid [gB, gH, gW] is walked in
 * 128-pixel bricks, input pixel = output pixel + (dx, dy, db) in A's coordinates with zero fill outside, which covers
 * stride-1 convs directly, stride-2 convs over a space-to-depth copy (b200lmd_space_to_depth_f16: sub-image
 * (py*2+px)*B+b) and transposed (dgrad) convs via the output mapping out_row = ((b*OH + y*sy+oy)*OW + x*sx+ox).
 * W is fp16 [N][wtaps][Cin].  mode: 0 row-major out (bias, chan_add[B,N], residual, fp16/fp32 out, accumulate),
 * 1 GEGLU, 2 head-split slabs.  Replaces every nn.Conv2d / nn.Linear of the UNet (models/unet_2d_condition.py:289,
 * 567; diffusers resnet.py conv1/conv2/conv_shortcut/Downsample2D/Upsample2D) and, with pre-transposed weights, the
 * activation gradients autograd computes for them (models/pipelines.py:56). */
typedef struct {
  const void* A; int aB, aH, aW, a_ld, Cin;
  const void* W; int N, wtaps;
  int gB, gH, gW;
  int ntaps; short taps[9][4];   /* dx, dy, db, wtap */
  int OH, OW, sy, sx, oy, ox;
  int mode; float alpha;
  const void* bias; const void* chan_add; int rows_per_img;
  const void* residual; int ldr;
  void* out; int ldo; void* out_f32; int ldo32; int accumulate; void* pre;
  int heads, head_dim, which0;
  void* rm[3]; int rm_alloc[3]; void* tr[3]; int tr_alloc[3];
} b200lmd_gemm_desc;
int b200lmd_gemm(const b200lmd_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------ layout / schedule helpers
 * HBM-bound, vectorised; each replaces a torch op on the reference path:
 *   copy_cols   torch.cat(dim=1) of skip connections (models/unet_2d_blocks.py:648,767) and its split adjoint
 *   copy_rows   torch.cat([x, objs], dim=1) / [:, :n_visual] of the GLIGEN fuser (models/attention.py:50) + gating
 *   upsample2x  F.interpolate(scale_factor=2, nearest) in diffusers Upsample2D, and its adjoint
 *   pack_latents / unpack_grad   NCHW fp32 latents <-> NHWC-8 fp16 activations (CFG duplicates with rep=2,
 *                                models/pipelines.py:420)
 *   timestep_embed  diffusers Timesteps(flip_sin_to_cos=True, freq_shift=0); silu_f32_to_f16 the MLP activation
 *   cfg_ddim_blend  eps_u + s(eps_c - eps_u), DDIM eta=0 step, frozen blend (models/pipelines.py:436-446)
 *   latent_update   z -= sqrt(1 - alpha_bar_t) * grad for still-active images (models/pipelines.py:62-69) */
int b200lmd_copy_cols_f16(const void* src, int ld_src, int src_off, void* dst, int ld_dst, int dst_off, long long rows,
                          int ncols, int accumulate, void* stream);
int b200lmd_copy_rows_f16(const void* src, int rows_per_img_src, int src_row0, void* dst, int rows_per_img_dst,
                          int dst_row0, int B, int nrows, int C, float alpha, int accumulate, void* stream);
int b200lmd_upsample2x_f16(const void* x, void* y, int B, int H, int W, int C, void* stream);
int b200lmd_upsample2x_bwd_f16(const void* dy, void* dx, int B, int H, int W, int C, int accumulate, void* stream);
int b200lmd_space_to_depth_f16(const void* x, void* y, int B, int Hout, int Wout, int C, void* stream);
int b200lmd_pack_latents(const void* z_f32, void* y_f16, int B, int Cz, int HW, int rep, void* stream);
int b200lmd_unpack_grad(const void* g_f32, int ld, void* out_f32, int B, int Cz, int HW, float scale, void* stream);
int b200lmd_timestep_embed(const void* t_f32, void* y_f16, int B, int dim, void* stream);
int b200lmd_silu_f32_to_f16(const void* x, void* y, long long n, void* stream);
int b200lmd_cfg_ddim_blend(void* z, const void* eps, int ld_eps, int B, int Cz, int HW, float guidance_scale,
                           float sa_t, float sb_t, float sa_p, float sb_p, int v_pred, const void* frozen,
                           const void* mask, void* stream);
int b200lmd_latent_update(void* z, const void* grad, int ld_g, int B, int Cz, int HW, float step_scale,
                          float inv_gscale, const int* active, void* stream);
/* Per-image guidance-loop predicate on the device (models/pipelines.py:30: `while loss/loss_scale > loss_threshold and
 * iteration < max_iter`, batched): `begin` resets the per-step counters and evaluates the predicate on the carried loss;
 * `advance` (after the guidance launch sequence and latent_update) reduces the loss partials parts[n_keys][B*heads] in a
 * fixed order, lets the images that were active take the new loss / count an iteration, records the trace row `slot`
 * and re-evaluates the predicate.  active[B] feeds b200lmd_latent_update; any_active[1] is what the host reads (and
 * only when the iteration count is data dependent).  loss / trace_loss are float64. */
int b200lmd_guidance_loop_begin(void* loss_f64, int* it, int* active, const int* has_boxes, int* any_active, int B,
                                double loss_scale, double threshold, int max_iter, void* stream);
int b200lmd_guidance_loop_advance(void* loss_f64, int* it, int* active, const int* has_boxes, void* trace_loss_f64,
                                  int* trace_active, int* any_active, const void* parts_f32, int n_keys, int B, int heads,
                                  double loss_scale, double threshold, int max_iter, int slot, void* stream);
/* Device-side latent composition between the two phases (utils/latents.py:37-118 compose_latents_with_alignment):
 * out[s,b,c,y,x] gathers from the per-box trajectory lat[s, owner-1, c, y-dy, x-dx] (zero outside) that owns the cell;
 * unowned cells are 0 for s > 0 and, for s = 0, the box-mask layer (bowner) over the background latent bg[B,C,H,W].
 * owner / bowner int32 [B,H,W] (1 + box index, 0 = none; the integer ownership maps are built on the host from the
 * masks, like the loss tables); shift int32 [BA,2] = (dx, dy) cells.  lat fp32 [S,BA,C,H,W], out fp32 [S,B,C,H,W]. */
int b200lmd_compose_latents(const void* lat, const void* bg, const int* owner, const int* bowner, const int* shift,
                            void* out, int S, int BA, int B, int C, int H, int W, void* stream);
/* VAE decoder helpers (models/pipelines.py:117-127; diffusers 0.18 AutoencoderKL.decode - the convolutions, GroupNorms
 * and the mid-block attention run on the conv / GN / GEMM entry points above):
 *   vae_prepare_latents  post_quant_conv(z * inv_scale) (1x1, 4 -> 4) from fp32 NCHW latents to fp16 NHWC-8
 *   softmax_rows         row softmax of fp32 scores [rows, n] -> fp16 probabilities (the single-head 512-wide mid-block
 *                        attention is two GEMMs around it)
 *   vae_to_uint8         round(clamp(x / 2 + 0.5, 0, 1) * 255) of the first 3 channels -> uint8 [pixels, 3] */
int b200lmd_vae_prepare_latents(const void* z_f32, const void* pq_w, const void* pq_b, void* y_f16, int B, int HW,
                                float inv_scale, void* stream);
int b200lmd_softmax_rows(const void* scores_f32, void* probs_f16, long long rows, int n, void* stream);
int b200lmd_vae_to_uint8(const void* x_f32, int ld, void* y_u8, long long pixels, void* stream);
/* GLIGEN PositionNet front end (models/unet_2d_condition.py:63-114): Fourier box features + phrase embeddings blended
 * with the learned null features by the object mask -> fp16 [rows, Demb+64] (input of PositionNet.linears[0]). */
int b200lmd_position_embed(const void* boxes, const void* masks, const void* emb, const void* null_pos,
                           const void* null_xyxy, void* out_f16, int rows, int Demb, void* stream);
int b200lmd_attn_delta_slab(const void* dO_slab, const void* o_tok, int ld_o, void* delta, int B, int heads, int nq,
                            int q_alloc, int head_dim, void* stream);

/* ------------------------------------------------------------------------------------------------ attention
 * Head-split projection: [M, nproj*C] = x[M,K] . W[nproj*C, K]^T scattered into per-head operand slabs
 * (to_q / to_k / to_v of models/attention_processor.py:426-438 plus head_to_batch_dim :190-199 in one pass):
 *   q  [B*heads, q_alloc, dp]  (projection index 0)      dp  = head_dim rounded up to 64, zero padded
 *   k  [B*heads, k_alloc, dp]  (projection index 1)
 *   vt [B*heads, d16, v_alloc] (projection index 2, stored d-major = V^T), d16 = b200lmd_round_d16(head_dim)
 * `which0` is the projection index of column 0 of W (so K/V-only or Q-only projections reuse the call).
 * Slabs must be zero-initialised once (padding is never written). rows_per_img = tokens per image in x. */
int b200lmd_round_dp(int head_dim);
int b200lmd_round_d16(int head_dim);
int b200lmd_project_heads_f16(const void* x, int ldx, const void* w, int M, int N, int K, int rows_per_img, int heads,
                              int head_dim, int which0, void* q, int q_alloc, void* k, int k_alloc, void* vt,
                              int v_alloc, void* stream);

/* General form: for projection i in {0,1,2} (relative to which0) write a row-major slab rm[i] [B*heads, rm_alloc[i],
 * dp] and/or a transposed slab tr[i] [B*heads, d16, tr_alloc[i]] (NULL = skip).  The backward kernels consume
 * row-major Q,K,V,dO and transposed Q^T,K^T,dO^T. */
int b200lmd_project_heads2_f16(const void* x, int ldx, const void* w, int M, int N, int K, int rows_per_img, int heads,
                               int head_dim, int which0, void* const* rm, const int* rm_alloc, void* const* tr,
                               const int* tr_alloc, void* stream);

/* Attention backward (activations only; weights are frozen): replaces torch.autograd through the attention of
 * models/attention_processor.py:216-233,447 inside torch.autograd.grad at models/pipelines.py:56.
 *   dq [B*nq, ld_dq], dk/dv [B*nk_store, ld] (NULL = not needed, e.g. text K/V are constants)
 *   delta: fp32 scratch [B*heads, q_alloc] (= rowsum(dO*O), computed here from the token-major o_tok/do_tok);
 *          pass NULL with a single KV tile to have it formed in-kernel (needed when dp_extra is given)
 *   dp_extra: optional fp32 [B*heads, nq, ext_ld] added to dP (the guidance loss's d loss / d P)
 *   dO may be NULL (no downstream gradient: last guidance layer), then only dp_extra drives dq. */
int b200lmd_attention_bwd_f16(const void* q, const void* k, const void* v, const void* dO, const void* qt,
                              const void* kt, const void* dOt, const void* lse2, void* delta, const void* o_tok,
                              int ld_o, const void* do_tok, int ld_do, const void* dp_extra, int ext_ld, void* dq,
                              int ld_dq, void* dk, int ld_dk, void* dv, int ld_dv, int nk_store, int B, int heads,
                              int nq, int nk, int q_alloc, int k_alloc, int head_dim, float scale, void* stream);

/* softmax(scale * q k^T) v over the slabs above -> out[B*nq, ldo] (head h at columns h*d..), lse2 (optional, fp32
 * [B*heads, q_alloc], log2-domain row statistic kept for the backward).  Replaces F.scaled_dot_product_attention at
 * models/attention_processor.py:355 and the baddbmm/softmax/bmm path :216-233,447 when no map is requested. */
int b200lmd_attention_fwd_f16(const void* q, const void* k, const void* vt, void* out, int ldo, void* lse2, int B,
                              int heads, int nq, int nk, int q_alloc, int k_alloc, int head_dim, float scale,
                              void* stream);

/* ------------------------------------------------------------------------------------------------ fused cross-attention + loss
 * The kernel the north star grades: cross-attention (T <= 128 text keys) with the LMD/LMD+ guidance loss evaluated in
 * the same launch.  Replaces AttnProcessor.__call__'s explicit path (models/attention_processor.py:407-483: probs,
 * P.V, save_attn_to_dict / return_token_ca_only) TOGETHER WITH utils/guidance.py:91-286 (compute_ca_lossv3) and the
 * d loss / d P half of torch.autograd.grad (models/pipelines.py:56).
 * Loss tables (device memory, built by the host mirror llm-groundeddiffusion_b200/guidance.py):
 *   terms[img_term_off[b] .. img_term_off[b+1]) for image b; weights already contain loss_scale and the normalisers
 *   1/(n_tokens * n_objects * n_keys) (utils/guidance.py:146,270) resp. w_ref/(n_boxes*n_tok*n_obj*n_keys*heads)
 *   (:233-237,284); masks are the scale_proportion rasters (utils/utils.py:57-70) as bytes [n_masks][n].
 * Outputs: loss_part[B*heads] (sum over heads and keys = loss * loss_scale), dp_extra[B*heads][n][ext_ld] =
 * gscale * d loss / d P (fed to b200lmd_attention_bwd_f16), optional maps. counters must be zero on first use. */
typedef struct {
  int type;            /* 0: fg/bg top-k energy, 1: reference-attention L1, 2: ratio-based energy (weight in w_fg) */
  int slot, mask, k_fg, k_bg;
  float w_fg, w_bg, w_ref;
  int ref;
} b200lmd_loss_term;
typedef struct {
  const int* img_term_off;
  const b200lmd_loss_term* terms;
  const unsigned char* masks;
  const float* refs;      /* [n_refs][heads][n] */
  const int* slot_tok;    /* [B][b200lmd_max_loss_slots()] token index per saved column, -1 terminated */
  float* pcol;            /* scratch [B*heads][slots][n] */
  int* counters;          /* [B*heads] */
  float* loss_part;       /* [B*heads] */
  float* dp_extra;        /* [B*heads][n][ext_ld] */
  int ext_ld;
  float gscale;
  float eps;              /* 1e-5 (utils/guidance.py:150) */
} b200lmd_xattn_loss;
int b200lmd_max_loss_slots(void);
int b200lmd_xattn_fwd_f16(const void* q, const void* k, const void* vt, void* out, int ldo, void* lse2, void* probs,
                          const int* save_tok, void* probs_tok, const b200lmd_xattn_loss* loss, int B, int heads,
                          int nq, int nk, int q_alloc, int k_alloc, int head_dim, float scale, void* stream);

/* The same op in ONE launch, projections included (the kernel BASELINE.json's roofline target names):
 *   out = residual + bias_o + softmax(scale (x Wq^T) K^T) V Wo^T      x, residual, out: fp16 [B*n, heads*head_dim]
 * Clusters of 8 CTAs (one per head) share each 128-row tile of x through TMA multicast; O is exchanged through L2
 * between the attention core and the output projection (o_scratch [B*n, C] fp16).  Requires heads == 8,
 * n % 128 == 0, head_dim in {64, 80, 160}, <= 80 text keys in 80-row K / V^T slabs (b200lmd_xattn_fused_supported).
 * q_slab (optional, [B*heads, n, dp]) and lse2 ([B*heads, n]) are what b200lmd_attention_bwd_f16 needs.
 * Replaces attn.to_q + get_attention_scores + bmm + to_out[0] (models/attention_processor.py:426-451) + the
 * residual add of BasicTransformerBlock (models/attention.py:220) + utils/guidance.py:91-286. */
int b200lmd_xattn_fused_supported(int heads, int head_dim, int n);
int b200lmd_xattn_fused_f16(const void* x, const void* wq, const void* k_slab, const void* vt_slab, const void* wo,
                            const void* bias_o, const void* residual, void* out, void* o_scratch, void* q_slab,
                            void* lse2, void* probs, const int* save_tok, void* probs_tok,
                            const b200lmd_xattn_loss* loss, int B, int n, int heads, int head_dim, int nk, int k_alloc,
                            float scale, void* stream);

/* ------------------------------------------------------------------------------------------------ BoxDiff loss
 * utils/boxdiff.py:20-161 (compute_ca_loss_boxdiff -> add_ca_loss_per_attn_map_to_loss_boxdiff ->
 * _compute_max_attention_per_index + _compute_loss) for B images from the fp16 maps [B*heads, n, T] the guidance forward
 * saved for each guidance key: key/head average, softmax(100 x) over tokens 1..T-2, optional 3x3 Gaussian (reflect
 * padding), inner/outer-box top-k means (k = floor(count*P), no clamp), corner (projection) L1; loss[b] (unscaled) and
 * out_scale * d loss / dP written to every key's dP_extra buffer (the same [n, T] matrix for all keys and heads). */
typedef struct {
  int tok;             /* token index of the phrase token in the prompt (1 .. T-2) */
  int mask;            /* row of masks[n_masks][n]: union-of-boxes cell mask of the phrase */
  int k_fg, k_bg;
  int corner;          /* row of corner[n_corner][2*side]: corner_x[side] then corner_y[side] (boxdiff.py:62-65) */
} b200lmd_boxdiff_term;
typedef struct {
  const void* maps[8];        /* fp16 [B*heads, n, T] per key */
  void* dp_extra[8];          /* fp32 [B*heads, n, ext_ld] per key */
  int n_keys, heads, n, side, T, ext_ld;
  const int* img_term_off;    /* [B+1] */
  const b200lmd_boxdiff_term* terms;
  const unsigned char* masks;
  const unsigned char* corner;
  float* mean;                /* scratch fp32 [B, n, T] */
  float* dA;                  /* scratch fp32 [B, n, T] */
  float* loss;                /* [B] */
  float kern[9];              /* 3x3 smoothing kernel, utils/attn.py:89-115 */
  int smooth;
  float out_scale;
} b200lmd_boxdiff;
int b200lmd_boxdiff_loss(const b200lmd_boxdiff* p, int B, void* stream);

/* The fused kernel's cross-CTA hand-shake counters live in one fixed-size per-device buffer that is allocated on first
 * use (outside stream capture), zero on entry to every launch and never freed (its address is baked into captured CUDA
 * graphs).  After a faulted launch call this to re-zero it (host-synchronous). */
int b200lmd_xattn_fused_reset(void);

/* profiling aid: launches an EMPTY kernel with the fused kernel's launch configuration (ctas CTAs in clusters of 4, 192
 * threads, the dynamic shared memory of the head_dim instantiation); bench.py times it for the launch + drain floor */
int b200lmd_xattn_fused_launch_floor(int head_dim, int ctas, void* stream);

/* profiling aid: device buffer [grid][8] of %globaltimer stamps written by the fused kernel (NULL disables) */
int b200lmd_set_debug_buffer(void* p);

/* ------------------------------------------------------------------------------------------------ normalisation
 * GroupNorm (+ optional SiLU) over NHWC fp16 x[B, n, C]: stats then apply (torch.nn.GroupNorm inside diffusers
 * ResnetBlock2D and models/transformer_2d.py:146,283).  sums: fp32 [B, groups, 2] (sum, sum of squares) per group,
 * written by the call (kept for the backward).  The reduction runs in a fixed order through a per-device scratch that
 * the library allocates on first use (outside stream capture): results repeat bit for bit; launches that share a
 * device must be ordered on one stream, like every other entry point here. */
int b200lmd_groupnorm_f16(const void* x, const void* gamma, const void* beta, void* y, void* sums, int B, int n, int C,
                          int groups, float eps, int silu, void* stream);
int b200lmd_groupnorm_bwd_f16(const void* dy, const void* x, const void* sums, const void* gamma, const void* beta,
                              void* dx, void* bsums, int B, int n, int C, int groups, float eps, int silu,
                              int accumulate, void* stream);
/* LayerNorm over the last dim of x[rows, C] (models/attention.py:114,133,149); stats fp32 [rows, 2] or NULL. */
int b200lmd_layernorm_f16(const void* x, const void* gamma, const void* beta, void* y, void* stats, long long rows,
                          int C, float eps, void* stream);
int b200lmd_layernorm_bwd_f16(const void* dy, const void* x, const void* stats, const void* gamma, void* dx,
                              long long rows, int C, int accumulate, void* stream);
int b200lmd_geglu_bwd_f16(const void* pre, const void* dy, void* dpre, long long rows, int F, void* stream);

#ifdef __cplusplus
}
#endif
#endif
