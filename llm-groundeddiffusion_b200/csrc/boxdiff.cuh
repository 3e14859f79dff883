// BoxDiff layout loss and its gradient on the GPU, sm_100a.
//
// Restates utils/boxdiff.py:20-117 (_compute_max_attention_per_index, _compute_loss) and :120-161
// (compute_ca_loss_boxdiff) of the reference for B images at once, from the fp16 cross-attention maps that the guidance
// forward saved ([B*heads, n, T] per guidance key, n = side*side <= 1024):
//   A[q,t]   = mean over (keys x heads) of P_key[h,q,t]                                   boxdiff.py:146
//   S[q,t']  = softmax_t'(100 * A[q, 1:T-1])                                              boxdiff.py:34-36
//   per phrase token: image = S[:, tok-1] (optionally 3x3 Gaussian, reflect padding)      boxdiff.py:44-79
//     fg = mean top-k(image * M), bg = mean top-k(image * (1-M)), k = floor(count * P)    boxdiff.py:81-89
//     dx = mean_x(|max_y image - proj_x M| * corner_x), dy likewise                       boxdiff.py:91-99
//   loss = sum max(0, 1 - fg) + sum max(0, bg) + sum dx + sum dy                          boxdiff.py:107-110
// and writes d(loss * scale)/dP - the same [n, T] matrix for every key and head, 1/(keys*heads) of d loss / dA - into
// the dP_extra buffers that attn_bwd_dq_kernel adds to dP (no autograd graph, no n x T map on the host).
//
// HBM-bound helper kernels (40 fp16 maps of n x T are read once, <= 1.6 MB per image):
//   boxdiff_mean_kernel   grid (B, n/ROWS): key/head average, fp32 [B, n, T]
//   boxdiff_loss_kernel   one CTA per image, one thread per map cell: softmax, per-token terms, gradient to d loss / dA
//   boxdiff_scatter_kernel  broadcast d loss / dA to the dP_extra buffers of all keys and heads
#pragma once
#include "ptx.cuh"

namespace b200 {

static constexpr int kBoxdiffMaxKeys = 8;
static constexpr int kBoxdiffMaxT = 80;

struct BoxdiffTerm {      // one (phrase, token) pair of one image
  int tok;                // token index in the prompt (1 <= tok <= T-2)
  int mask;               // row of `masks` (union-of-boxes cell mask of the phrase) [n]
  int k_fg, k_bg;         // floor(count * P) - no clamp (boxdiff.py:82,87)
  int corner;             // row of `corner` tables: [2*side] = corner_x[side] then corner_y[side]
};

struct BoxdiffParams {
  const __half* maps[kBoxdiffMaxKeys];   // [B*heads, n, T] each
  float* dp_extra[kBoxdiffMaxKeys];      // [B*heads, n, ext_ld] each
  int n_keys, heads, n, side, T, ext_ld;
  const int* img_term_off;               // [B+1]
  const BoxdiffTerm* terms;
  const uint8_t* masks;                  // [n_masks][n]
  const uint8_t* corner;                 // [n_corner][2*side]
  float* mean;                           // scratch [B, n, T] fp32
  float* dA;                             // scratch [B, n, T] fp32: d loss / d A (unscaled)
  float* loss;                           // [B]
  float kern[9];                         // 3x3 smoothing kernel (row-major), utils/attn.py:89-115
  int smooth;
  float out_scale;                       // gscale * loss_scale / (n_keys * heads)
};

__global__ void boxdiff_mean_kernel(const __grid_constant__ BoxdiffParams p) {
  const int b = blockIdx.y;
  const long long per_img = (long long)p.n * p.T;
  const float inv = 1.f / (float)(p.n_keys * p.heads);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_img;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < p.n_keys; ++k) {
      const __half* base = p.maps[k] + (long long)b * p.heads * per_img + i;
      for (int h = 0; h < p.heads; ++h) acc += __half2float(base[(long long)h * per_img]);
    }
    p.mean[(long long)b * per_img + i] = acc * inv;
  }
}

// block-wide helpers for up to 1024 threads
__device__ __forceinline__ float bd_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
__device__ __forceinline__ int bd_block_count(bool pred) { return __syncthreads_count(pred ? 1 : 0); }

// mean of the k largest values of v over the block's cells (each thread holds one value, v >= 0) and the selection
// (greater than the k-th value, plus the lowest-index ties).  Bisection on the float bit pattern: 31 counting passes.
__device__ __forceinline__ float bd_topk_mean(float v, int k, bool valid, bool& selected, float* red, int* ired) {
  const uint32_t bits = valid ? __float_as_uint(v) : 0u;
  uint32_t lo = 0u, hi = 0x7F800000u;           // count(bits >= lo) >= k, count(bits >= hi) < k
  while (hi - lo > 1u) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const int c = bd_block_count(valid && bits >= mid);
    if (c >= k) lo = mid; else hi = mid;
  }
  const bool gt = valid && bits > lo, tie = valid && bits == lo;
  const int c_gt = bd_block_count(gt);
  const float s_gt = bd_block_sum(gt ? v : 0.f, red);
  // lowest-index ties first: exclusive prefix count of ties over thread index
  const unsigned bal = __ballot_sync(0xffffffffu, tie);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) ired[warp] = __popc(bal);
  __syncthreads();
  int before = __popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) before += ired[w];
  const int need = k - c_gt;
  selected = gt || (tie && before < need);
  return (s_gt + (float)need * __uint_as_float(lo)) / (float)k;
}

// one CTA per image, blockDim.x == n (one thread per map cell, row-major y*side + x)
__global__ void boxdiff_loss_kernel(const __grid_constant__ BoxdiffParams p) {
  extern __shared__ float bd_smem[];
  const int b = blockIdx.x, q = threadIdx.x, n = p.n, side = p.side, T = p.T;
  const int y = q / side, x = q - y * side;
  float* img = bd_smem;                  // [n] current (smoothed) image
  float* gimg = img + n;                 // [n] gradient w.r.t. the smoothed image
  float* colmax = gimg + n;              // [side] max over y per column, then [side] max over x per row
  float* red = colmax + 2 * side;        // [32]
  int* ired = reinterpret_cast<int*>(red + 32);   // [32]
  int* amax = ired + 32;                 // [2*side] argmax cell index per column / row
  const float* A = p.mean + ((long long)b * n + q) * T;
  // softmax over tokens 1 .. T-2 of 100 * A (row statistics kept in registers; probabilities recomputed when needed)
  float m = -INFINITY;
  for (int t = 1; t < T - 1; ++t) m = fmaxf(m, 100.f * A[t]);
  float l = 0.f;
  for (int t = 1; t < T - 1; ++t) l += expf(100.f * A[t] - m);
  const float inv_l = 1.f / l;
  float gdot = 0.f;                      // sum_t g[q,t] S[q,t] (softmax backward)
  float loss_acc = 0.f;                  // thread 0 only
  float* dA = p.dA + ((long long)b * n + q) * T;
  for (int t = 0; t < T; ++t) dA[t] = 0.f;            // accumulates g[q,t] * S[q,t] first
  const int t0 = p.img_term_off[b], t1 = p.img_term_off[b + 1];
  for (int ti = t0; ti < t1; ++ti) {
    const BoxdiffTerm term = p.terms[ti];
    const float s_raw = expf(100.f * A[term.tok] - m) * inv_l;       // S[q, tok-1]
    __syncthreads();
    img[q] = s_raw;
    __syncthreads();
    float v = s_raw;
    if (p.smooth) {                       // 3x3 correlation on the reflect-padded image (F.pad mode='reflect')
      v = 0.f;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          int yy = y + dy, xx = x + dx;
          yy = yy < 0 ? -yy : (yy >= side ? 2 * side - 2 - yy : yy);
          xx = xx < 0 ? -xx : (xx >= side ? 2 * side - 2 - xx : xx);
          v += p.kern[(dy + 1) * 3 + (dx + 1)] * img[yy * side + xx];
        }
    }
    const bool in_mask = p.masks[(long long)term.mask * n + q] != 0;
    const uint8_t* cm = p.corner + (long long)term.corner * 2 * side;
    float g = 0.f;
    // inner / outer box constraints
    bool sel;
    const float fg = bd_topk_mean(in_mask ? v : 0.f, term.k_fg, true, sel, red, ired);
    if (1.f - fg > 0.f && sel && in_mask) g += -1.f / (float)term.k_fg;          // max(0, 1 - fg)
    const float bg = bd_topk_mean(in_mask ? 0.f : v, term.k_bg, true, sel, red, ired);
    if (bg > 0.f && sel && !in_mask) g += 1.f / (float)term.k_bg;                 // max(0, bg)
    // corner constraint: projections of the image (max over y per column, max over x per row) vs those of the mask
    __syncthreads();
    img[q] = v;
    gimg[q] = p.masks[(long long)term.mask * n + q] ? 1.f : 0.f;
    __syncthreads();
    float dist = 0.f;
    if (q < 2 * side) {
      const bool col = q < side;
      const int idx = col ? q : q - side;
      float best = -INFINITY, pm = 0.f;
      int arg = 0;
      for (int j = 0; j < side; ++j) {
        const int cell = col ? j * side + idx : idx * side + j;
        if (img[cell] > best) { best = img[cell]; arg = cell; }      // first maximum, like torch.max
        pm = fmaxf(pm, gimg[cell]);
      }
      const float d = best - pm;
      const float w = cm[q] ? 1.f / (float)side : 0.f;
      dist = fabsf(d) * w;
      colmax[q] = w * ((d > 0.f) - (d < 0.f));
      amax[q] = arg;
    }
    const float dsum = bd_block_sum(dist, red);
    __syncthreads();
    for (int j = 0; j < 2 * side; ++j)
      if (amax[j] == q) g += colmax[j];
    if (q == 0) loss_acc += fmaxf(0.f, 1.f - fg) + fmaxf(0.f, bg) + dsum;
    // back through the smoothing (adjoint of the reflect-padded correlation)
    float gs = g;
    if (p.smooth) {
      __syncthreads();
      gimg[q] = g;
      __syncthreads();
      // cell q receives, from every output cell c in its 3x3 neighbourhood, the kernel weights of the taps of c that
      // land on q (directly, or through the reflect padding at the border)
      gs = 0.f;
      for (int cy = y - 1; cy <= y + 1; ++cy)
        for (int cx = x - 1; cx <= x + 1; ++cx) {
          if (cy < 0 || cy >= side || cx < 0 || cx >= side) continue;
          float wsum = 0.f;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              int yy = cy + dy, xx = cx + dx;
              yy = yy < 0 ? -yy : (yy >= side ? 2 * side - 2 - yy : yy);
              xx = xx < 0 ? -xx : (xx >= side ? 2 * side - 2 - xx : xx);
              if (yy == y && xx == x) wsum += p.kern[(dy + 1) * 3 + (dx + 1)];
            }
          gs += wsum * gimg[cy * side + cx];
        }
    }
    // softmax backward needs g[q,t] S[q,t] per token and their sum
    dA[term.tok] += gs * s_raw;
    gdot += gs * s_raw;
  }
  // d loss / dA[q,t] = 100 * (g S - S * sum_t' g S) for t in 1..T-2, zero for the first and last token
  for (int t = 1; t < T - 1; ++t) {
    const float s = expf(100.f * A[t] - m) * inv_l;
    dA[t] = 100.f * (dA[t] - s * gdot);
  }
  if (q == 0) p.loss[b] = loss_acc;
}

__global__ void boxdiff_scatter_kernel(const __grid_constant__ BoxdiffParams p, int B) {
  const long long per_img = (long long)p.n * p.T;
  const long long total = (long long)B * p.heads * per_img;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % p.T);
    const long long r = i / p.T;                 // (b*heads + h) * n + q
    const int q = (int)(r % p.n);
    const long long bh = r / p.n;
    const int b = (int)(bh / p.heads);
    const float g = p.dA[((long long)b * p.n + q) * p.T + t] * p.out_scale;
    for (int k = 0; k < p.n_keys; ++k) p.dp_extra[k][(bh * p.n + q) * p.ext_ld + t] = g;
  }
}

}  // namespace b200
