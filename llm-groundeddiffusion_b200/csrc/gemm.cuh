// Implicit-GEMM family on tcgen05 (sm_100a).
//
//   D[m, n] = sum_{tap, c}  A[pixel(m) + off(tap), c] * W[n, tap, c]          (fp16 in, fp32 accumulate in TMEM)
//
// One kernel covers every dense contraction on the UNet path: Linear / 1x1 conv (1 tap, a "pixel" is a token),
// conv3x3 (9 taps, TMA zero-fill supplies the padding), stride-2 conv and its transposed dgrad (tap tables over
// space-to-depth sub-images), and all dgrads (same kernel, pre-transposed weights).
//   A : NHWC activations viewed through a 4-D tensor map (C, W, H, B); an M tile is a (tb x th x tw) pixel brick
//   W : [N][taps][C] viewed through a 3-D tensor map (C, taps, N)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue (TMEM -> regs -> HBM).
#pragma once
#include "ptx.cuh"

namespace b200 {

enum EpiMode : int {
  EPI_ROWMAJOR = 0,  // out[row, n] fp16 (optionally also fp32)
  EPI_GEGLU = 1,     // tile columns [0,64) = value, [64,128) = gate (weights interleaved by the loader)
  EPI_HEADS = 2,     // scatter into per-head Q / K (row-major, padded) and V^T (d-major) slabs
};

struct GemmTap {
  int16_t dx, dy, db, wtap;
};

struct GemmParams {
  // ---- A side / tiling
  int B, H, W;          // logical output-pixel grid the M tiles walk over (for plain GEMM: B=1,H=1,W=M)
  int tb, th, tw;       // brick shape, tb*th*tw == 128
  int tiles_x, tiles_y, tiles_b;
  int N;                // output columns
  int Cin;              // contraction length per tap
  int ntaps;
  GemmTap taps[9];
  // ---- output pixel mapping: out_row = ((b*OH + y*sy+oy)*OW + x*sx+ox)
  int OH, OW, sy, sx, oy, ox;
  // ---- epilogue
  int mode;
  float alpha;
  const float* bias;       // [N] or null
  const float* chan_add;   // [B_img, N] added per image (time-embedding projection) or null
  int rows_per_img;        // rows (pixels/tokens) per image, for chan_add and EPI_HEADS
  const __half* residual;  // [rows, ldr] or null (added after everything else)
  int ldr;
  __half* out;             // fp16 output or null
  int ldo;
  float* out_f32;          // optional fp32 output (same indexing, ld = ldo32)
  int ldo32;
  int accumulate_out;      // out += result (fp16 read-modify-write), used by gradient accumulation
  __half* pre;             // EPI_GEGLU: optional pre-activation dump [rows, 2*ldo] in tile-interleaved order
  // ---- EPI_HEADS
  int C;                   // channels per projection (heads*d)
  int heads, d;
  int which0;              // projection index of column 0 (0 = Q, 1 = K, 2 = V)
  int dp, d16;             // padded head dims of the row-major / transposed slabs
  __half* rm[3]; int rm_alloc[3];   // row-major slab per projection  [B*heads, rm_alloc, dp]   (or null)
  __half* tr[3]; int tr_alloc[3];   // transposed slab per projection [B*heads, d16, tr_alloc]  (or null)
};

// exact (erf) GELU, F.gelu default (models/attention.py:329).  erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7,
// far below fp16 output resolution): two SFU ops + ~10 FMAs instead of libdevice erff's ~40 instructions - the GEGLU
// epilogue is issue-bound on the short-K feed-forward GEMMs.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = poly * t * ex2_approx(-z * z * 1.4426950408889634f);   // 1 - erf(z)
  const float erf_abs = 1.f - e;
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int BLOCK_K = 64;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y;
  const int xt = m_tile % p.tiles_x;
  const int yt = (m_tile / p.tiles_x) % p.tiles_y;
  const int bt = m_tile / (p.tiles_x * p.tiles_y);
  const int x0 = xt * p.tw, y0 = yt * p.th, b0 = bt * p.tb;
  const int n0 = n_tile * BLOCK_N;
  const int cpb = (p.Cin + 63) >> 6;
  const int num_kb = cpb * p.ntaps;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<BLOCK_N>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / cpb;
        const int cc = kb - tap * cpb;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        const GemmTap t = p.taps[tap];
        tma_load_4d(sa, &tmA, &full_bar[stage], cc * 64, x0 + t.dx, y0 + t.dy, b0 + t.db);
        tma_load_3d(sb, &tmB, &full_bar[stage], cc * 64, t.wtap, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_f16(128, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        const uint64_t adesc = make_desc_k_sw128(sa);
        const uint64_t bdesc = make_desc_k_sw128(sb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_f16_ss(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);
        if (kb == num_kb - 1) tc_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====================
    const int quad = warp & 3;
    const int r = quad * 32 + lane_id();  // row inside the tile == TMEM lane
    const int xl = r % p.tw;
    const int yl = (r / p.tw) % p.th;
    const int bl = r / (p.tw * p.th);
    const int x = x0 + xl, y = y0 + yl, b = b0 + bl;
    const bool row_ok = (x < p.W) && (y < p.H) && (b < p.B);
    const long long orow = ((long long)b * p.OH + (y * p.sy + p.oy)) * p.OW + (x * p.sx + p.ox);
    const int img = p.rows_per_img > 0 ? (int)(orow / p.rows_per_img) : 0;

    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);

    if (p.mode == EPI_ROWMAJOR || p.mode == EPI_HEADS) {
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        if (n0 + c0 >= p.N) break;  // uniform across the warp
        uint32_t v[32];
        tmem_ld_x32(taddr + c0, v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + c0 + g * 8;
          if (n >= p.N) break;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = __uint_as_float(v[g * 8 + j]) * p.alpha;
            const int nn = n + j;
            if (nn < p.N) {
              if (p.bias) a += __ldg(p.bias + nn);
              if (p.chan_add) a += __ldg(p.chan_add + (long long)img * p.N + nn);
            }
            f[j] = a;
          }
          if (p.mode == EPI_ROWMAJOR) {
            const bool full = (n + 8 <= p.N);
            if (p.residual) {
              if (full && (p.ldr % 8 == 0)) {
                uint4 rr = *reinterpret_cast<const uint4*>(p.residual + orow * p.ldr + n);
                const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 t2 = __half22float2(rh[j]);
                  f[2 * j] += t2.x;
                  f[2 * j + 1] += t2.y;
                }
              } else {
                for (int j = 0; j < 8 && n + j < p.N; ++j) f[j] += __half2float(p.residual[orow * p.ldr + n + j]);
              }
            }
            if (p.out_f32) {
              for (int j = 0; j < 8 && n + j < p.N; ++j) {
                float* o = p.out_f32 + orow * p.ldo32 + n + j;
                *o = p.accumulate_out ? (*o + f[j]) : f[j];
              }
            }
            if (p.out) {
              if (full && (p.ldo % 8 == 0)) {
                uint4* dst = reinterpret_cast<uint4*>(p.out + orow * p.ldo + n);
                if (p.accumulate_out) {
                  uint4 old = *dst;
                  const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float2 t2 = __half22float2(oh[j]);
                    f[2 * j] += t2.x;
                    f[2 * j + 1] += t2.y;
                  }
                }
                uint4 st;
                st.x = pack_h2(f[0], f[1]);
                st.y = pack_h2(f[2], f[3]);
                st.z = pack_h2(f[4], f[5]);
                st.w = pack_h2(f[6], f[7]);
                *dst = st;
              } else {
                for (int j = 0; j < 8 && n + j < p.N; ++j) {
                  __half* o = p.out + orow * p.ldo + n + j;
                  float a = f[j];
                  if (p.accumulate_out) a += __half2float(*o);
                  *o = __float2half_rn(a);
                }
              }
            }
          } else {  // EPI_HEADS: 8-column groups never straddle a head (d % 8 == 0)
            const int which = p.which0 + n / p.C;
            const int cc = n % p.C;
            const int head = cc / p.d;
            const int j0 = cc % p.d;
            const int tok = (int)(orow % p.rows_per_img);
            const long long bh = (long long)img * p.heads + head;
            if (p.tr[which]) {
              __half* dst = p.tr[which] + (bh * p.d16 + j0) * (long long)p.tr_alloc[which] + tok;
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[(long long)j * p.tr_alloc[which]] = __float2half_rn(f[j]);
            }
            if (p.rm[which]) {
              uint4 st;
              st.x = pack_h2(f[0], f[1]);
              st.y = pack_h2(f[2], f[3]);
              st.z = pack_h2(f[4], f[5]);
              st.w = pack_h2(f[6], f[7]);
              *reinterpret_cast<uint4*>(p.rm[which] + (bh * p.rm_alloc[which] + tok) * (long long)p.dp + j0) = st;
            }
          }
        }
      }
    } else if (p.mode == EPI_GEGLU) {
      // BLOCK_N == 128: value columns [0,64), gate columns [64,128); output columns n_tile*64 + [0,64)
      if constexpr (BLOCK_N == 128) {
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t vv[16], gg[16];
          tmem_ld_x16(taddr + c0, vv);
          tmem_ld_x16(taddr + 64 + c0, gg);
          tmem_ld_wait();
          if (!row_ok) continue;
          float o[16];
          float pv[16], pg[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(vv[j]);
            float g = __uint_as_float(gg[j]);
            if (p.bias) {
              a += __ldg(p.bias + n0 + c0 + j);
              g += __ldg(p.bias + n0 + 64 + c0 + j);
            }
            pv[j] = a;
            pg[j] = g;
            o[j] = a * gelu_erf(g);
          }
          const int on = n_tile * 64 + c0;
          uint4* dst = reinterpret_cast<uint4*>(p.out + orow * p.ldo + on);
          uint4 s0, s1;
          s0.x = pack_h2(o[0], o[1]); s0.y = pack_h2(o[2], o[3]); s0.z = pack_h2(o[4], o[5]); s0.w = pack_h2(o[6], o[7]);
          s1.x = pack_h2(o[8], o[9]); s1.y = pack_h2(o[10], o[11]); s1.z = pack_h2(o[12], o[13]); s1.w = pack_h2(o[14], o[15]);
          dst[0] = s0;
          dst[1] = s1;
          if (p.pre) {
            __half* pr = p.pre + orow * (long long)(2 * p.ldo) + n0 + c0;
            uint4* d0 = reinterpret_cast<uint4*>(pr);
            uint4* d1 = reinterpret_cast<uint4*>(pr + 64);
            uint4 t;
            t.x = pack_h2(pv[0], pv[1]); t.y = pack_h2(pv[2], pv[3]); t.z = pack_h2(pv[4], pv[5]); t.w = pack_h2(pv[6], pv[7]);
            d0[0] = t;
            t.x = pack_h2(pv[8], pv[9]); t.y = pack_h2(pv[10], pv[11]); t.z = pack_h2(pv[12], pv[13]); t.w = pack_h2(pv[14], pv[15]);
            d0[1] = t;
            t.x = pack_h2(pg[0], pg[1]); t.y = pack_h2(pg[2], pg[3]); t.z = pack_h2(pg[4], pg[5]); t.w = pack_h2(pg[6], pg[7]);
            d1[0] = t;
            t.x = pack_h2(pg[8], pg[9]); t.y = pack_h2(pg[10], pg[11]); t.z = pack_h2(pg[12], pg[13]); t.w = pack_h2(pg[14], pg[15]);
            d1[1] = t;
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BLOCK_N>(tmem_base);
  }
}

}  // namespace b200
