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

#ifdef __cplusplus
}
#endif
#endif
