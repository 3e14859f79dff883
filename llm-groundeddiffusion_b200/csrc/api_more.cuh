// C-ABI: descriptor-driven implicit GEMM (every conv / linear / dgrad variant) and the HBM-bound helper kernels.
// Included by api_ops.cu.
#pragma once

extern "C" int b200lmd_gemm(const b200lmd_gemm_desc* d, void* stream) {
  return b200::guarded([&] {
    using namespace b200;
    GemmBuild b;
    b.A = (const __half*)d->A; b.aB = d->aB; b.aH = d->aH; b.aW = d->aW; b.a_ld = d->a_ld; b.Cin = d->Cin;
    b.Wt = (const __half*)d->W; b.N = d->N; b.wtaps = d->wtaps;
    b.gB = d->gB; b.gH = d->gH; b.gW = d->gW;
    b.ntaps = d->ntaps;
    if (d->ntaps < 1 || d->ntaps > 9) throw std::runtime_error("gemm: ntaps out of range");
    for (int i = 0; i < d->ntaps; ++i) b.taps[i] = GemmTap{d->taps[i][0], d->taps[i][1], d->taps[i][2], d->taps[i][3]};
    GemmParams ep = default_epilogue();
    ep.OH = d->OH; ep.OW = d->OW; ep.sy = d->sy; ep.sx = d->sx; ep.oy = d->oy; ep.ox = d->ox;
    ep.mode = d->mode; ep.alpha = d->alpha;
    ep.bias = (const float*)d->bias; ep.chan_add = (const float*)d->chan_add; ep.rows_per_img = d->rows_per_img;
    ep.residual = (const __half*)d->residual; ep.ldr = d->ldr;
    ep.out = (__half*)d->out; ep.ldo = d->ldo; ep.out_f32 = (float*)d->out_f32; ep.ldo32 = d->ldo32;
    ep.accumulate_out = d->accumulate; ep.pre = (__half*)d->pre;
    if (d->mode == EPI_HEADS) {
      if (d->head_dim % 8) throw std::runtime_error("head_dim must be a multiple of 8");
      ep.C = d->heads * d->head_dim; ep.heads = d->heads; ep.d = d->head_dim; ep.which0 = d->which0;
      ep.dp = round_dp(d->head_dim); ep.d16 = round_d16(d->head_dim);
      for (int i = 0; i < 3; ++i) {
        ep.rm[i] = (__half*)d->rm[i]; ep.rm_alloc[i] = d->rm_alloc[i];
        ep.tr[i] = (__half*)d->tr[i]; ep.tr_alloc[i] = d->tr_alloc[i];
      }
    }
    if (d->Cin % 8 || d->a_ld % 8) throw std::runtime_error("gemm: channel counts must be multiples of 8");
    run_gemm(build_gemm(b, ep), (cudaStream_t)stream);
  });
}

#define B200_EW(call)                                   \
  return b200::guarded([&] {                            \
    using namespace b200;                               \
    cudaStream_t st = (cudaStream_t)stream;             \
    call;                                               \
    B200_CHECK(cudaGetLastError());                     \
  })

extern "C" int b200lmd_copy_cols_f16(const void* src, int ld_src, int src_off, void* dst, int ld_dst, int dst_off,
                                     long long rows, int ncols, int accumulate, void* stream) {
  B200_EW((copy_cols_kernel<<<ew_grid(rows * (ncols / 8)), 256, 0, st>>>((const __half*)src, ld_src, src_off,
                                                                         (__half*)dst, ld_dst, dst_off, rows, ncols,
                                                                         accumulate)));
}

namespace b200 {
// dst[b*rd + dst_row0 + i, :] (+)= alpha * src[b*rs + src_row0 + i, :]  for i < nrows   (fp16, C % 8 == 0)
__global__ void copy_rows_kernel(const __half* __restrict__ src, int rs, int src_row0, __half* __restrict__ dst, int rd,
                                 int dst_row0, int B, int nrows, int C, float alpha, int accumulate) {
  const int vecs = C >> 3;
  const long long total = (long long)B * nrows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long r = i / vecs;
    const int b = (int)(r / nrows), k = (int)(r % nrows);
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(src + ((long long)b * rs + src_row0 + k) * C + v * 8), f);
    uint4* d = reinterpret_cast<uint4*>(dst + ((long long)b * rd + dst_row0 + k) * C + v * 8);
    if (accumulate) {
      float o[8];
      unpack8(*d, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = o[j] + alpha * f[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= alpha;
    }
    *d = pack8(f);
  }
}
// grad fp32 NHWC [B, HW, ld] -> fp32 NCHW [B, Cz, HW] scaled
__global__ void unpack_grad_kernel(const float* __restrict__ g, int ld, float* __restrict__ out, int B, int Cz, int HW,
                                   float scale) {
  const long long total = (long long)B * Cz * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % Cz);
    const int b = (int)(i / ((long long)HW * Cz));
    out[i] = g[((long long)b * HW + p) * ld + c] * scale;
  }
}
// delta[bh, q] from the row-major dO slab [BH, nq_alloc, dp] and token-major O [B*nq, ld]
__global__ void attn_delta_slab_kernel(const __half* __restrict__ dO, int dp, const __half* __restrict__ O, int ld_o,
                                       float* __restrict__ delta, int B, int heads, int nq, int nq_alloc, int d) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * heads * nq;
  for (long long w = blockIdx.x * (long long)warps + (threadIdx.x >> 5); w < total; w += (long long)gridDim.x * warps) {
    const int q = (int)(w % nq);
    const long long bh = w / nq;
    const int b = (int)(bh / heads), h = (int)(bh % heads);
    float acc = 0.f;
    for (int j = lane; j < d; j += 32)
      acc += __half2float(dO[(bh * nq_alloc + q) * dp + j]) * __half2float(O[((long long)b * nq + q) * ld_o + h * d + j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) delta[bh * nq_alloc + q] = acc;
  }
}
}  // namespace b200

extern "C" int b200lmd_copy_rows_f16(const void* src, int rows_per_img_src, int src_row0, void* dst,
                                     int rows_per_img_dst, int dst_row0, int B, int nrows, int C, float alpha,
                                     int accumulate, void* stream) {
  B200_EW((copy_rows_kernel<<<ew_grid((long long)B * nrows * (C / 8)), 256, 0, st>>>(
      (const __half*)src, rows_per_img_src, src_row0, (__half*)dst, rows_per_img_dst, dst_row0, B, nrows, C, alpha,
      accumulate)));
}
extern "C" int b200lmd_upsample2x_f16(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  B200_EW((upsample2x_kernel<<<ew_grid((long long)B * 4 * H * W * (C / 8)), 256, 0, st>>>((const __half*)x, (__half*)y,
                                                                                         B, H, W, C)));
}
extern "C" int b200lmd_upsample2x_bwd_f16(const void* dy, void* dx, int B, int H, int W, int C, int accumulate,
                                          void* stream) {
  B200_EW((upsample2x_bwd_kernel<<<ew_grid((long long)B * H * W * (C / 8)), 256, 0, st>>>(
      (const __half*)dy, (__half*)dx, B, H, W, C, accumulate)));
}
extern "C" int b200lmd_space_to_depth_f16(const void* x, void* y, int B, int Hout, int Wout, int C, void* stream) {
  B200_EW((space_to_depth_kernel<<<ew_grid((long long)4 * B * Hout * Wout * (C / 8)), 256, 0, st>>>(
      (const __half*)x, (__half*)y, B, Hout, Wout, C)));
}
extern "C" int b200lmd_pack_latents(const void* z_f32, void* y_f16, int B, int Cz, int HW, int rep, void* stream) {
  B200_EW((pack_latents_kernel<<<ew_grid((long long)B * rep * HW), 256, 0, st>>>((const float*)z_f32, (__half*)y_f16, B,
                                                                                Cz, HW, rep)));
}
extern "C" int b200lmd_unpack_grad(const void* g_f32, int ld, void* out_f32, int B, int Cz, int HW, float scale,
                                   void* stream) {
  B200_EW((unpack_grad_kernel<<<ew_grid((long long)B * Cz * HW), 256, 0, st>>>((const float*)g_f32, ld, (float*)out_f32,
                                                                              B, Cz, HW, scale)));
}
extern "C" int b200lmd_timestep_embed(const void* t_f32, void* y_f16, int B, int dim, void* stream) {
  B200_EW((timestep_embed_kernel<<<(B * (dim / 2) + 255) / 256, 256, 0, st>>>((const float*)t_f32, (__half*)y_f16, B,
                                                                             dim)));
}
extern "C" int b200lmd_silu_f32_to_f16(const void* x, void* y, long long n, void* stream) {
  B200_EW((silu_f32_to_f16_kernel<<<ew_grid(n), 256, 0, st>>>((const float*)x, (__half*)y, n)));
}
extern "C" int b200lmd_cfg_ddim_blend(void* z, const void* eps, int ld_eps, int B, int Cz, int HW,
                                      float guidance_scale, float sa_t, float sb_t, float sa_p, float sb_p,
                                      int v_pred, const void* frozen, const void* mask, void* stream) {
  B200_EW((cfg_ddim_blend_kernel<<<ew_grid((long long)B * Cz * HW), 256, 0, st>>>(
      (float*)z, (const float*)eps, ld_eps, B, Cz, HW, guidance_scale, sa_t, sb_t, sa_p, sb_p, v_pred,
      (const float*)frozen, (const float*)mask)));
}
extern "C" int b200lmd_latent_update(void* z, const void* grad, int ld_g, int B, int Cz, int HW, float step_scale,
                                     float inv_gscale, const int* active, void* stream) {
  B200_EW((latent_update_kernel<<<ew_grid((long long)B * Cz * HW), 256, 0, st>>>(
      (float*)z, (const float*)grad, ld_g, B, Cz, HW, step_scale, inv_gscale, active)));
}
extern "C" int b200lmd_attn_delta_slab(const void* dO_slab, const void* o_tok, int ld_o, void* delta, int B, int heads,
                                       int nq, int q_alloc, int head_dim, void* stream) {
  B200_EW((attn_delta_slab_kernel<<<ew_grid((long long)B * heads * nq * 32), 256, 0, st>>>(
      (const __half*)dO_slab, b200::round_dp(head_dim), (const __half*)o_tok, ld_o, (float*)delta, B, heads, nq, q_alloc,
      head_dim)));
}

namespace b200 {
// GLIGEN PositionNet front end (models/unet_2d_condition.py:63-114): Fourier features of the boxes (8 frequencies,
// temperature 100, layout [freq][sin|cos][xyxy]) and phrase embeddings, each blended with its learned null feature by
// the per-object mask; output fp16 [B*N, Demb + 64] = the input of PositionNet.linears[0].
__global__ void position_embed_kernel(const float* __restrict__ boxes, const float* __restrict__ masks,
                                      const float* __restrict__ emb, const float* __restrict__ null_pos,
                                      const float* __restrict__ null_xyxy, __half* __restrict__ out, int rows, int Demb) {
  const int ld = Demb + 64;
  const long long total = (long long)rows * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i % ld);
    const float m = masks[r];
    float v;
    if (c < Demb) {
      v = emb[(long long)r * Demb + c] * m + (1.f - m) * null_pos[c];
    } else {
      const int k = c - Demb;
      const int f = k >> 3, sc = (k >> 2) & 1, coord = k & 3;
      const float a = boxes[r * 4 + coord] * powf(100.f, (float)f / 8.f);
      v = (sc ? cosf(a) : sinf(a)) * m + (1.f - m) * null_xyxy[k];
    }
    out[i] = __float2half_rn(v);
  }
}
}  // namespace b200

extern "C" int b200lmd_position_embed(const void* boxes, const void* masks, const void* emb, const void* null_pos,
                                      const void* null_xyxy, void* out_f16, int rows, int Demb, void* stream) {
  B200_EW((position_embed_kernel<<<ew_grid((long long)rows * (Demb + 64)), 256, 0, st>>>(
      (const float*)boxes, (const float*)masks, (const float*)emb, (const float*)null_pos, (const float*)null_xyxy,
      (__half*)out_f16, rows, Demb)));
}

namespace b200 {
// Per-image guidance-loop state on the device (models/pipelines.py:16-82 batched over B images): after every guidance
// iteration the loss partials the attention kernels wrote are reduced in a fixed order, the images that were active take
// the new loss and count one more iteration, and the loop predicate `loss / loss_scale > threshold && it < max_iter` is
// re-evaluated per image - so the host never uploads an active mask and reads back one small trace row, and only when
// the iteration count is data dependent.  One block, one thread per image.
struct GuidanceLoopState {
  double* loss;          // [B] scaled loss carried across steps (the reference's `loss`, initialised to 10000 * ... on the host)
  int* it;               // [B] iterations done in the current step
  int* active;           // [B] 1 = this image runs the next iteration (read by latent_update_kernel)
  const int* has_boxes;  // [B]
  double* trace_loss;    // [cap][B] loss after each iteration of the current step
  int* trace_active;     // [cap][B] which images took part in it
  int* any_active;       // [1]
};
__global__ void guidance_loop_begin_kernel(GuidanceLoopState s, int B, double loss_scale, double threshold, int max_iter) {
  __shared__ int any;
  if (threadIdx.x == 0) any = 0;
  __syncthreads();
  const int b = threadIdx.x;
  if (b < B) {
    s.it[b] = 0;
    const int a = s.has_boxes[b] && (s.loss[b] / loss_scale > threshold) && (0 < max_iter);
    s.active[b] = a;
    if (a) atomicOr(&any, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) *s.any_active = any;
}
__global__ void guidance_loop_advance_kernel(GuidanceLoopState s, const float* __restrict__ parts, int n_keys, int B,
                                             int heads, double loss_scale, double threshold, int max_iter, int slot) {
  __shared__ int any;
  if (threadIdx.x == 0) any = 0;
  __syncthreads();
  const int b = threadIdx.x;
  if (b < B) {
    const int was = s.active[b];
    if (was) {
      double acc = 0.0;                       // fixed order: key-major, then head
      for (int k = 0; k < n_keys; ++k)
        for (int h = 0; h < heads; ++h) acc += (double)parts[((long long)k * B + b) * heads + h];
      s.loss[b] = acc;
      s.it[b] += 1;
    }
    s.trace_loss[(long long)slot * B + b] = s.loss[b];
    s.trace_active[(long long)slot * B + b] = was;
    const int a = was && (s.loss[b] / loss_scale > threshold) && (s.it[b] < max_iter);
    s.active[b] = a;
    if (a) atomicOr(&any, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) *s.any_active = any;
}
}  // namespace b200

extern "C" int b200lmd_guidance_loop_begin(void* loss_f64, int* it, int* active, const int* has_boxes, int* any_active,
                                           int B, double loss_scale, double threshold, int max_iter, void* stream) {
  return b200::guarded([&] {
    using namespace b200;
    if (B > 1024) throw std::runtime_error("guidance loop state: at most 1024 images per batch");
    GuidanceLoopState s{(double*)loss_f64, it, active, has_boxes, nullptr, nullptr, any_active};
    guidance_loop_begin_kernel<<<1, ((B + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(s, B, loss_scale, threshold, max_iter);
    B200_CHECK(cudaGetLastError());
  });
}
extern "C" int b200lmd_guidance_loop_advance(void* loss_f64, int* it, int* active, const int* has_boxes,
                                             void* trace_loss_f64, int* trace_active, int* any_active, const void* parts_f32,
                                             int n_keys, int B, int heads, double loss_scale, double threshold,
                                             int max_iter, int slot, void* stream) {
  return b200::guarded([&] {
    using namespace b200;
    if (B > 1024) throw std::runtime_error("guidance loop state: at most 1024 images per batch");
    GuidanceLoopState s{(double*)loss_f64, it, active, has_boxes, (double*)trace_loss_f64, trace_active, any_active};
    guidance_loop_advance_kernel<<<1, ((B + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(
        s, (const float*)parts_f32, n_keys, B, heads, loss_scale, threshold, max_iter, slot);
    B200_CHECK(cudaGetLastError());
  });
}

namespace b200 {
// Device-side latent composition (utils/latents.py:37-83 compose_latents, after the per-box shifts of :85-118): every
// output cell gathers from the per-box trajectory that owns it.
//   lat     fp32 [S, BA, C, H, W]  all steps of the per-box generations (stays on the GPU after Phase A)
//   owner   int32 [B, H, W]        1 + batch index (into BA) of the LAST box in composition order whose shifted mask
//                                  covers the cell, 0 = background                       (latents.py:70-78)
//   bowner  int32 [B, H, W]        same for the enlarged box masks that blend step 0 into the background (:62-68)
//   shift   int32 [BA, 2]          integer (dx, dy) cell shift of each box: out[y, x] = lat[y - dy, x - dx], zero fill
//   out     fp32 [S, B, C, H, W]   composed; step 0: bg -> box-mask layer -> mask layer, steps > 0: mask layer over zeros
__global__ void compose_latents_kernel(const float* __restrict__ lat, const float* __restrict__ bg,
                                       const int* __restrict__ owner, const int* __restrict__ bowner,
                                       const int* __restrict__ shift, float* __restrict__ out, int S, int BA, int B, int C,
                                       int H, int W) {
  const long long total = (long long)S * B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int c = (int)((i / ((long long)W * H)) % C);
    const int b = (int)((i / ((long long)W * H * C)) % B);
    const int s = (int)(i / ((long long)W * H * C * B));
    const int cell = (b * H + y) * W + x;
    auto fetch = [&](int box) -> float {
      const int sx = x - shift[2 * box], sy = y - shift[2 * box + 1];
      if (sx < 0 || sx >= W || sy < 0 || sy >= H) return 0.f;
      return lat[((((long long)s * BA + box) * C + c) * H + sy) * W + sx];
    };
    float v = 0.f;
    const int o = owner[cell];
    if (o > 0) v = fetch(o - 1);
    else if (s == 0) {
      const int bo = bowner[cell];
      v = bo > 0 ? fetch(bo - 1) : bg[(((long long)b * C + c) * H + y) * W + x];
    }
    out[i] = v;
  }
}
}  // namespace b200

extern "C" int b200lmd_compose_latents(const void* lat, const void* bg, const int* owner, const int* bowner,
                                       const int* shift, void* out, int S, int BA, int B, int C, int H, int W,
                                       void* stream) {
  B200_EW((compose_latents_kernel<<<ew_grid((long long)S * B * C * H * W), 256, 0, st>>>(
      (const float*)lat, (const float*)bg, owner, bowner, shift, (float*)out, S, BA, B, C, H, W)));
}

namespace b200 {
// ---- VAE decoder helpers (models/pipelines.py:117-127 decode; diffusers 0.18 AutoencoderKL.decode)
// latents fp32 NCHW [B, 4, HW] -> post_quant_conv (1x1, 4 -> 4, bias) of z * inv_scale -> fp16 NHWC-8 [B, HW, 8]
__global__ void vae_prepare_latents_kernel(const float* __restrict__ z, const float* __restrict__ w /*[4][4]*/,
                                           const float* __restrict__ bias /*[4]*/, __half* __restrict__ y, int B, int HW,
                                           float inv_scale) {
  const long long total = (long long)B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = z[(b * 4 + c) * HW + p] * inv_scale;
    float f[8];
#pragma unroll
    for (int o = 0; o < 4; ++o) f[o] = bias[o] + w[o * 4] * v[0] + w[o * 4 + 1] * v[1] + w[o * 4 + 2] * v[2] + w[o * 4 + 3] * v[3];
#pragma unroll
    for (int o = 4; o < 8; ++o) f[o] = 0.f;
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(f);
  }
}
// row softmax: fp32 scores [rows, n] (already scaled) -> fp16 probabilities [rows, n]; one block per row
__global__ void softmax_rows_kernel(const float* __restrict__ s, __half* __restrict__ pr, long long rows, int n) {
  __shared__ float red[32];
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* row = reinterpret_cast<const float4*>(s + r * n);
    const int n4 = n >> 2;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 f = row[i];
      m = fmaxf(fmaxf(m, fmaxf(f.x, f.y)), fmaxf(f.z, f.w));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    float l = 0.f;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 f = row[i];
      l += __expf(f.x - m) + __expf(f.y - m) + __expf(f.z - m) + __expf(f.w - m);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
    __syncthreads();
    l = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) l += red[w];
    const float inv = 1.f / l;
    __half2* out = reinterpret_cast<__half2*>(pr + r * n);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 f = row[i];
      out[2 * i] = __floats2half2_rn(__expf(f.x - m) * inv, __expf(f.y - m) * inv);
      out[2 * i + 1] = __floats2half2_rn(__expf(f.z - m) * inv, __expf(f.w - m) * inv);
    }
    __syncthreads();
  }
}
// decoder output fp32 NHWC [B*HW, ld] (first 3 channels) -> uint8 [B*HW, 3] = round(clamp(x / 2 + 0.5, 0, 1) * 255)
__global__ void vae_to_uint8_kernel(const float* __restrict__ x, int ld, unsigned char* __restrict__ y, long long pixels) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels * 3;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 3;
    const int c = (int)(i - p * 3);
    const float v = fminf(fmaxf(x[p * ld + c] * 0.5f + 0.5f, 0.f), 1.f);
    y[i] = (unsigned char)rintf(v * 255.f);
  }
}
}  // namespace b200

extern "C" int b200lmd_vae_prepare_latents(const void* z_f32, const void* pq_w, const void* pq_b, void* y_f16, int B,
                                           int HW, float inv_scale, void* stream) {
  B200_EW((vae_prepare_latents_kernel<<<ew_grid((long long)B * HW), 256, 0, st>>>(
      (const float*)z_f32, (const float*)pq_w, (const float*)pq_b, (__half*)y_f16, B, HW, inv_scale)));
}
extern "C" int b200lmd_softmax_rows(const void* scores_f32, void* probs_f16, long long rows, int n, void* stream) {
  return b200::guarded([&] {
    using namespace b200;
    if (n % 8) throw std::runtime_error("softmax_rows: n must be a multiple of 8");
    long long blocks = rows < (long long)kNumSMs * 16 ? rows : (long long)kNumSMs * 16;
    softmax_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)scores_f32,
                                                                             (__half*)probs_f16, rows, n);
    B200_CHECK(cudaGetLastError());
  });
}
extern "C" int b200lmd_vae_to_uint8(const void* x_f32, int ld, void* y_u8, long long pixels, void* stream) {
  B200_EW((vae_to_uint8_kernel<<<ew_grid(pixels * 3), 256, 0, st>>>((const float*)x_f32, ld, (unsigned char*)y_u8,
                                                                    pixels)));
}

// runtime switches (kept for A/B measurements of kernel generations; defaults are the fastest correct variants)
extern "C" int b200lmd_set_option(const char* name, int value) {
  return b200::guarded([&] {
    std::string n(name);
    if (n == "attn_v2") b200::attn_use_v2() = value != 0;
    else if (n == "gemm_v2") b200::gemm_use_v2() = value != 0;
    else if (n == "gemm_v3") b200::gemm_use_v3() = value != 0;
    else if (n == "fused_loss_stage") b200::fused_loss_stage() = value != 0;
    else throw std::runtime_error("unknown option " + n);
  });
}
