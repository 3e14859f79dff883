// HBM-bound kernels of the UNet path: GroupNorm(+SiLU) / LayerNorm forward+backward, GEGLU backward, layout
// shuffles (concat, nearest-2x upsample and its adjoint, space-to-depth), latent pack/unpack, timestep sinusoid and
// the fused CFG + DDIM + frozen-blend update.  All are coalesced, 16-byte vectorised over the channel dimension
// (activations are [rows, C] fp16 with C % 8 == 0) and launched with grids that are multiples of the SM count.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

static constexpr int kNumSMs = 148;

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
  __half2 c = __floats2half2_rn(f[4], f[5]), d = __floats2half2_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c);
  u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// stats[b, g] = (sum, sumsq) over n rows x cpg channels.  Bit-reproducible from run to run: no floating-point
// atomics anywhere - per-thread channel sums go to shared memory, one thread per (group, moment) folds them in a fixed
// order, the block partial is stored to part[b][chunk][group][2], and the last block of the image to finish (ticket on
// counter[b], self-resetting) adds the chunks in index order into sums[b].  Nothing needs zeroing before the launch.
// grid = (chunks, B); each block walks rows [chunk*rows_per, ...) of image b with threads over 8-channel vectors.
// dynamic shared memory: 2 * rif * C floats (rif = rows in flight per block).
__device__ __forceinline__ void gn_fold_and_publish(float* sch, int rifC, int rif, int C, int cpg, int groups,
                                                    float* __restrict__ part, int* __restrict__ counter,
                                                    float* __restrict__ out) {
  __shared__ int s_last;
  const int b = blockIdx.y, G2 = groups * 2, chunks = gridDim.x;
  for (int i = threadIdx.x; i < G2; i += blockDim.x) {
    const int g = i >> 1;
    const float* src = sch + (i & 1) * rifC + g * cpg;
    float acc = 0.f;
    for (int tr = 0; tr < rif; ++tr)
      for (int c = 0; c < cpg; ++c) acc += src[tr * C + c];
    part[((long long)b * chunks + blockIdx.x) * G2 + i] = acc;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&counter[b], 1) == chunks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // four threads per output, each over every fourth chunk (loads in flight together), combined in a fixed order
  for (int idx = threadIdx.x; idx < G2 * 4; idx += blockDim.x) {
    const int i = idx >> 2;
    float acc = 0.f;
#pragma unroll 4
    for (int ch = idx & 3; ch < chunks; ch += 4) acc += __ldcg(&part[((long long)b * chunks + ch) * G2 + i]);
    sch[idx] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G2; i += blockDim.x)
    out[(long long)b * G2 + i] = ((sch[4 * i] + sch[4 * i + 1]) + sch[4 * i + 2]) + sch[4 * i + 3];
  if (threadIdx.x == 0) counter[b] = 0;
}

__global__ void gn_stats_kernel(const __half* __restrict__ x, float* __restrict__ sums, float* __restrict__ part,
                                int* __restrict__ counter, int n, int C, int groups, int rows_per) {
  extern __shared__ float sch[];  // [2][rif][C] per-thread channel sums
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(n, r0 + rows_per);
  // threads as (rows in flight) x (8-channel vectors of a row): every thread streams, loads of a warp are contiguous
  const int tpr = min(vecs, (int)blockDim.x);
  const int rif = blockDim.x / tpr;
  const int rifC = rif * C;
  const int tr = threadIdx.x / tpr, tv = threadIdx.x - tr * tpr;
  for (int v = tv; v < vecs && tr < rif; v += tpr) {
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    const __half* base = x + ((long long)b * n + r0 + tr) * C + v * 8;
    const long long step = (long long)rif * C;
#pragma unroll 4
    for (int r = r0 + tr; r < r1; r += rif, base += step) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(base), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] += f[i] * f[i];
      }
    }
    float* d = sch + tr * C + v * 8;
    *reinterpret_cast<float4*>(d) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(s[4], s[5], s[6], s[7]);
    *reinterpret_cast<float4*>(d + rifC) = make_float4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<float4*>(d + rifC + 4) = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  gn_fold_and_publish(sch, rifC, rif, C, cpg, groups, part, counter, sums);
}

// y = silu?( (x - mean) * rstd * gamma + beta ), fp16 out.  sums -> mean/rstd on the fly.
__global__ void gn_apply_kernel(const __half* __restrict__ x, const float* __restrict__ sums,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                __half* __restrict__ y, int B, int n, int C, int groups, float eps, int do_silu,
                                int rows_per) {
  // grid = (chunks, B).  Per-channel scale / shift of image b once per block (shared memory), then a pure
  // fma (+ silu) stream with the same (rows in flight) x (vectors of a row) thread layout as the statistics pass.
  extern __shared__ float s_ss[];   // [C] scale, [C] shift
  float* s_scale = s_ss;
  float* s_shift = s_ss + C;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const float inv_cnt = 1.f / ((float)n * cpg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = sums[((long long)b * groups + g) * 2] * inv_cnt;
    const float var = sums[((long long)b * groups + g) * 2 + 1] * inv_cnt - mean * mean;
    const float sc = rsqrtf(fmaxf(var, 0.f) + eps) * gamma[c];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - mean * sc;
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(n, r0 + rows_per);
  const int tpr = min(vecs, (int)blockDim.x);
  const int rif = blockDim.x / tpr;
  const int tr = threadIdx.x / tpr, tv = threadIdx.x - tr * tpr;
  for (int v = tv; v < vecs && tr < rif; v += tpr) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = s_scale[v * 8 + j];
      sh[j] = s_shift[v * 8 + j];
    }
    long long off = ((long long)b * n + r0 + tr) * C + v * 8;
    const long long step = (long long)rif * C;
#pragma unroll 4
    for (int r = r0 + tr; r < r1; r += rif, off += step) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + off), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(f[j], sc[j], sh[j]);
        f[j] = do_silu ? silu_f(t) : t;
      }
      *reinterpret_cast<uint4*>(y + off) = pack8(f);
    }
  }
}

// backward, pass 1: per (b, group) sums of  g1 = sum(dyh * gamma)  and  g2 = sum(dyh * gamma * xhat)
// where dyh = dy * silu'(pre) (pre = xhat*gamma+beta) when do_silu.
__global__ void gn_bwd_stats_kernel(const __half* __restrict__ dy, const __half* __restrict__ x,
                                    const float* __restrict__ sums, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float* __restrict__ bsums,
                                    float* __restrict__ part, int* __restrict__ counter, int n, int C, int groups,
                                    float eps, int do_silu, int rows_per) {
  extern __shared__ float sch[];  // [2][rif][C], reduced in a fixed order like gn_stats_kernel
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(n, r0 + rows_per);
  const float inv_cnt = 1.f / ((float)n * cpg);
  const int tpr = min(vecs, (int)blockDim.x);
  const int rif = blockDim.x / tpr;
  const int rifC = rif * C;
  const int tr = threadIdx.x / tpr, tv = threadIdx.x - tr * tpr;
  for (int v = tv; v < vecs && tr < rif; v += tpr) {
    float mean[8], rstd[8], ga[8], be[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const int g = c / cpg;
      mean[j] = sums[((long long)b * groups + g) * 2] * inv_cnt;
      const float var = sums[((long long)b * groups + g) * 2 + 1] * inv_cnt - mean[j] * mean[j];
      rstd[j] = rsqrtf(fmaxf(var, 0.f) + eps);
      ga[j] = gamma[c];
      be[j] = beta[c];
      s1[j] = s2[j] = 0.f;
    }
    const long long off0 = ((long long)b * n + r0) * C + v * 8;
#pragma unroll 2
    for (int r = r0 + tr; r < r1; r += rif) {
      float fx[8], fd[8];
      const long long off = off0 + (long long)(r - r0) * C;
      unpack8(*reinterpret_cast<const uint4*>(x + off), fx);
      unpack8(*reinterpret_cast<const uint4*>(dy + off), fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fx[j] - mean[j]) * rstd[j];
        float d = fd[j];
        if (do_silu) d *= silu_grad(xh * ga[j] + be[j]);
        d *= ga[j];
        s1[j] += d;
        s2[j] += d * xh;
      }
    }
    float* d = sch + tr * C + v * 8;
    *reinterpret_cast<float4*>(d) = make_float4(s1[0], s1[1], s1[2], s1[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
    *reinterpret_cast<float4*>(d + rifC) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    *reinterpret_cast<float4*>(d + rifC + 4) = make_float4(s2[4], s2[5], s2[6], s2[7]);
  }
  __syncthreads();
  gn_fold_and_publish(sch, rifC, rif, C, cpg, groups, part, counter, bsums);
}

// backward, pass 2: dx = rstd * (d - mean(d) - xhat * mean(d * xhat)),  d = dyh*gamma ; dx (+)= into out
__global__ void gn_bwd_apply_kernel(const __half* __restrict__ dy, const __half* __restrict__ x,
                                    const float* __restrict__ sums, const float* __restrict__ bsums,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    __half* __restrict__ dx, int B, int n, int C, int groups, float eps, int do_silu,
                                    int accumulate, int rows_per) {
  // grid = (chunks, B); per-channel constants live in registers across the row walk (same thread layout as the
  // statistics pass)
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const int vecs = C >> 3;
  const float inv_cnt = 1.f / ((float)n * cpg);
  const int r0 = blockIdx.x * rows_per;
  const int r1 = min(n, r0 + rows_per);
  const int tpr = min(vecs, (int)blockDim.x);
  const int rif = blockDim.x / tpr;
  const int tr = threadIdx.x / tpr, tv = threadIdx.x - tr * tpr;
  for (int v = tv; v < vecs && tr < rif; v += tpr) {
    float mean[8], rstd[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      const long long sg = ((long long)b * groups + c / cpg) * 2;
      mean[j] = sums[sg] * inv_cnt;
      const float var = sums[sg + 1] * inv_cnt - mean[j] * mean[j];
      rstd[j] = rsqrtf(fmaxf(var, 0.f) + eps);
      ga[j] = gamma[c];
      be[j] = beta[c];
      m1[j] = bsums[sg] * inv_cnt;
      m2[j] = bsums[sg + 1] * inv_cnt;
    }
    long long off = ((long long)b * n + r0 + tr) * C + v * 8;
    const long long step = (long long)rif * C;
#pragma unroll 2
    for (int r = r0 + tr; r < r1; r += rif, off += step) {
      float fx[8], fd[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(x + off), fx);
      unpack8(*reinterpret_cast<const uint4*>(dy + off), fd);
      if (accumulate) unpack8(*reinterpret_cast<const uint4*>(dx + off), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fx[j] - mean[j]) * rstd[j];
        float d = fd[j];
        if (do_silu) d *= silu_grad(xh * ga[j] + be[j]);
        d *= ga[j];
        const float rr = rstd[j] * (d - m1[j] - xh * m2[j]);
        o[j] = accumulate ? o[j] + rr : rr;
      }
      *reinterpret_cast<uint4*>(dx + off) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// one warp per row; stats[row] = (mean, rstd) kept for the backward
__global__ void ln_fwd_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, __half* __restrict__ y, float* __restrict__ stats,
                              long long rows, int C, float eps) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  for (long long row = blockIdx.x * (long long)warps + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * warps) {
    const __half* xr = x + row * C;
    float s = 0.f, q = 0.f;
    if (vecs <= 5 * 32) {
      // the row stays in registers between the statistics and the normalisation: one global read (C <= 1280)
      uint4 buf[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int v = lane + 32 * k;
        buf[k] = v < vecs ? *reinterpret_cast<const uint4*>(xr + v * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float f[8];
        unpack8(buf[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s += f[j];
          q += f[j] * f[j];
        }
      }
      s = warp_sum(s);
      q = warp_sum(q);
      const float mean = s / C;
      const float rstd = rsqrtf(fmaxf(q / C - mean * mean, 0.f) + eps);
      if (stats && lane == 0) {
        stats[row * 2] = mean;
        stats[row * 2 + 1] = rstd;
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int v = lane + 32 * k;
        if (v < vecs) {
          float f[8];
          unpack8(buf[k], f);
          const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * gg[j] + bb[j];
          *reinterpret_cast<uint4*>(y + row * C + v * 8) = pack8(f);
        }
      }
      continue;
    }
    for (int v = lane; v < vecs; v += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += f[j];
        q += f[j] * f[j];
      }
    }
    s = warp_sum(s);
    q = warp_sum(q);
    const float mean = s / C;
    const float rstd = rsqrtf(fmaxf(q / C - mean * mean, 0.f) + eps);
    if (stats && lane == 0) {
      stats[row * 2] = mean;
      stats[row * 2 + 1] = rstd;
    }
    for (int v = lane; v < vecs; v += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(xr + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * gamma[v * 8 + j] + beta[v * 8 + j];
      *reinterpret_cast<uint4*>(y + row * C + v * 8) = pack8(f);
    }
  }
}

// LayerNorm forward for narrow rows (C <= 768): one warp works on RPW rows at a time, so VPL * RPW 16-byte loads per
// lane are in flight instead of one or two (the one-row-per-warp kernel keeps ~15 KB per SM in flight at C = 320 and
// ran at 0.38 of the HBM rate, profiles/r2/norm_kernels_ncu.md).  Same arithmetic and statistics layout.
template <int VPL, int RPW, int MINB>
__global__ void __launch_bounds__(256, MINB) ln_fwd_rows_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, __half* __restrict__ y, float* __restrict__ stats,
                                   long long rows, int C, float eps) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  const float inv_c = 1.f / (float)C;
  for (long long row0 = (blockIdx.x * (long long)warps + (threadIdx.x >> 5)) * RPW; row0 < rows;
       row0 += (long long)gridDim.x * warps * RPW) {
    uint4 buf[RPW][VPL];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        buf[r][k] = (v < vecs && row0 + r < rows) ? *reinterpret_cast<const uint4*>(x + (row0 + r) * C + v * 8)
                                                  : make_uint4(0, 0, 0, 0);
      }
    float mean[RPW], rstd[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        float f[8];
        unpack8(buf[r][k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s += f[j];
          q += f[j] * f[j];
        }
      }
      s = warp_sum(s);
      q = warp_sum(q);
      mean[r] = s * inv_c;
      rstd[r] = rsqrtf(fmaxf(q * inv_c - mean[r] * mean[r], 0.f) + eps);
      if (stats && lane == 0 && row0 + r < rows) {
        stats[(row0 + r) * 2] = mean[r];
        stats[(row0 + r) * 2 + 1] = rstd[r];
      }
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = lane + 32 * k;
      if (v < vecs) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          if (row0 + r < rows) {
            float f[8];
            unpack8(buf[r][k], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean[r]) * rstd[r] * gg[j] + bb[j];
            *reinterpret_cast<uint4*>(y + (row0 + r) * C + v * 8) = pack8(f);
          }
        }
      }
    }
  }
}

__global__ void ln_bwd_kernel(const __half* __restrict__ dy, const __half* __restrict__ x,
                              const float* __restrict__ stats, const float* __restrict__ gamma,
                              __half* __restrict__ dx, long long rows, int C, int accumulate) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  for (long long row = blockIdx.x * (long long)warps + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * warps) {
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int v = lane; v < vecs; v += 32) {
      float fx[8], fd[8];
      unpack8(*reinterpret_cast<const uint4*>(x + row * C + v * 8), fx);
      unpack8(*reinterpret_cast<const uint4*>(dy + row * C + v * 8), fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = fd[j] * gamma[v * 8 + j];
        s1 += d;
        s2 += d * (fx[j] - mean) * rstd;
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    for (int v = lane; v < vecs; v += 32) {
      float fx[8], fd[8], o[8];
      unpack8(*reinterpret_cast<const uint4*>(x + row * C + v * 8), fx);
      unpack8(*reinterpret_cast<const uint4*>(dy + row * C + v * 8), fd);
      if (accumulate) unpack8(*reinterpret_cast<const uint4*>(dx + row * C + v * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fx[j] - mean) * rstd;
        const float r = rstd * (fd[j] * gamma[v * 8 + j] - s1 - xh * s2);
        o[j] = accumulate ? o[j] + r : r;
      }
      *reinterpret_cast<uint4*>(dx + row * C + v * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEGLU backward
// pre: [rows, 2F] tile-interleaved (value block 64 | gate block 64), dy: [rows, F] -> dpre same layout as pre
__global__ void geglu_bwd_kernel(const __half* __restrict__ pre, const __half* __restrict__ dy,
                                 __half* __restrict__ dpre, long long rows, int F) {
  const int vecs = F >> 3;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long row = i / vecs;
    const int col = v * 8;
    const int t = col >> 6, j = col & 63;
    const long long pv = row * 2 * F + t * 128 + j;
    float fv[8], fg[8], fd[8], ov[8], og[8];
    unpack8(*reinterpret_cast<const uint4*>(pre + pv), fv);
    unpack8(*reinterpret_cast<const uint4*>(pre + pv + 64), fg);
    unpack8(*reinterpret_cast<const uint4*>(dy + row * F + col), fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float g = fg[k];
      const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
      ov[k] = fd[k] * g * cdf;
      og[k] = fd[k] * fv[k] * (cdf + g * pdf);
    }
    *reinterpret_cast<uint4*>(dpre + pv) = pack8(ov);
    *reinterpret_cast<uint4*>(dpre + pv + 64) = pack8(og);
  }
}

// ------------------------------------------------------------------------------------------------ layout shuffles
// dst[row, dst_off : dst_off+Csrc] = src[row, :]   (concat along channels; also the "split" adjoint with accumulate)
__global__ void copy_cols_kernel(const __half* __restrict__ src, int ld_src, int src_off, __half* __restrict__ dst,
                                 int ld_dst, int dst_off, long long rows, int Ccopy, int accumulate) {
  const int vecs = Ccopy >> 3;
  const long long total = rows * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long row = i / vecs;
    uint4 s = *reinterpret_cast<const uint4*>(src + row * ld_src + src_off + v * 8);
    uint4* d = reinterpret_cast<uint4*>(dst + row * ld_dst + dst_off + v * 8);
    if (accumulate) {
      float a[8], b[8];
      unpack8(s, a);
      unpack8(*d, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
      *d = pack8(a);
    } else {
      *d = s;
    }
  }
}

// nearest 2x upsample: y[b, 2h+i, 2w+j, :] = x[b, h, w, :]
__global__ void upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W, int C) {
  const int vecs = C >> 3;
  const long long total = (long long)B * 4 * H * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int ox = (int)(p % (2 * W));
    p /= 2 * W;
    const int oy = (int)(p % (2 * H));
    const int b = (int)(p / (2 * H));
    const long long src = (((long long)b * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 8;
    *reinterpret_cast<uint4*>(y + (((long long)b * 2 * H + oy) * 2 * W + ox) * C + v * 8) =
        *reinterpret_cast<const uint4*>(x + src);
  }
}
// adjoint: dx[b,h,w,:] (+)= sum of the 4 upsampled cells
__global__ void upsample2x_bwd_kernel(const __half* __restrict__ dy, __half* __restrict__ dx, int B, int H, int W,
                                      int C, int accumulate) {
  const int vecs = C >> 3;
  const long long total = (long long)B * H * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    uint4* d = reinterpret_cast<uint4*>(dx + (((long long)b * H + y) * W + x) * C + v * 8);
    if (accumulate) unpack8(*d, acc);
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(dy + (((long long)b * 2 * H + 2 * y + dyy) * 2 * W + 2 * x + dxx) * C +
                                                v * 8),
                f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    *d = pack8(acc);
  }
}

// space-to-depth by parity: y[(py*2+px)*B + b, h, w, :] = x[b, 2h+py, 2w+px, :]   (H, W = output half-res)
__global__ void space_to_depth_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W,
                                      int C) {
  const int vecs = C >> 3;
  const long long total = (long long)4 * B * H * W * vecs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    p /= H;
    const int b = (int)(p % B);
    const int par = (int)(p / B);
    const int py = par >> 1, px = par & 1;
    *reinterpret_cast<uint4*>(y + i * 8) = *reinterpret_cast<const uint4*>(
        x + (((long long)b * 2 * H + 2 * h + py) * 2 * W + 2 * w + px) * C + v * 8);
  }
}

// ------------------------------------------------------------------------------------------------ latents / schedule
// z fp32 NCHW [B, 4, H, W] (optionally only channel-block `rep` copies: out batch = B*rep) -> fp16 NHWC with 8 chans
__global__ void pack_latents_kernel(const float* __restrict__ z, __half* __restrict__ y, int B, int Cz, int HW,
                                    int rep) {
  const long long total = (long long)B * rep * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int ob = (int)(i / HW);
    const int b = ob % B;  // batch layout of the CFG pass: [uncond copies ; cond copies]
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = (c < Cz) ? z[((long long)b * Cz + c) * HW + p] : 0.f;
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(f);
  }
}

// sinusoidal timestep embedding, [cos | sin] order (flip_sin_to_cos=True, freq_shift=0), fp16 [B, dim]
__global__ void timestep_embed_kernel(const float* __restrict__ t, __half* __restrict__ y, int B, int dim) {
  const int half_dim = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half_dim) return;
  const int b = i / half_dim, k = i % half_dim;
  const float freq = expf(-9.210340371976184f * (float)k / (float)half_dim);
  const float a = t[b] * freq;
  y[(long long)b * dim + k] = __float2half_rn(cosf(a));
  y[(long long)b * dim + half_dim + k] = __float2half_rn(sinf(a));
}

// y = silu(x) elementwise on fp32 -> fp16 (time-embedding MLP)
__global__ void silu_f32_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(silu_f(x[i]));
}

// CFG combine + DDIM(eta=0) step + frozen-latent blend, fp32 latents NCHW [B,4,HW]; eps is fp32 NHWC-8 [2B,HW,8]
// (rows [0,B) = uncond, [B,2B) = cond).  coef = {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)}.
__global__ void cfg_ddim_blend_kernel(float* __restrict__ z, const float* __restrict__ eps, int ld_eps, int B, int Cz,
                                      int HW, float guidance_scale, float sa_t, float sb_t, float sa_p, float sb_p,
                                      int v_pred, const float* __restrict__ frozen, const float* __restrict__ mask) {
  const long long total = (long long)B * Cz * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % Cz);
    const int b = (int)(i / ((long long)HW * Cz));
    const float eu = eps[((long long)b * HW + p) * ld_eps + c];
    const float ec = eps[((long long)(B + b) * HW + p) * ld_eps + c];
    const float m = eu + guidance_scale * (ec - eu);
    const float x = z[i];
    float x0, e;
    if (!v_pred) {
      x0 = (x - sb_t * m) / sa_t;
      e = m;
    } else {
      x0 = sa_t * x - sb_t * m;
      e = sa_t * m + sb_t * x;
    }
    float r = sa_p * x0 + sb_p * e;
    if (frozen) {
      const float mk = mask[(long long)b * HW + p];   // per-image mask [B, HW]
      r = frozen[i] * mk + r * (1.f - mk);
    }
    z[i] = r;
  }
}

// guidance latent update: z[b] -= step_scale * grad[b] * active[b]; grad is fp32 NHWC-8 [B, HW, 8] scaled by gscale
__global__ void latent_update_kernel(float* __restrict__ z, const float* __restrict__ grad, int ld_g, int B, int Cz,
                                     int HW, float step_scale, float inv_gscale, const int* __restrict__ active) {
  const long long total = (long long)B * Cz * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % Cz);
    const int b = (int)(i / ((long long)HW * Cz));
    if (active && !active[b]) continue;
    z[i] -= step_scale * inv_gscale * grad[((long long)b * HW + p) * ld_g + c];
  }
}

// y[i] = a*x[i] (fp16), used for tanh-gated residual pieces and misc
__global__ void fill_f32_kernel(float* __restrict__ p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

inline int ew_grid(long long work_items, int threads = 256) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace b200
