// Implicit GEMM, second generation: persistent CTAs, double-buffered TMEM accumulators, coalesced epilogue.
//
// Why (profiles/bench_kernels.py on B200, round 1): the first kernel's epilogue wrote 16 bytes per thread per row, i.e.
// 32 different cache lines per warp store; the LSU serialised them (~19 k cycles per 128x128 tile for the K=320 GEGLU
// projection = 160 TFLOP/s, an HBM-bound op running at 1/20 of its bandwidth bound).  Here
//   * each CTA loops over output tiles (grid = min(#tiles, #SMs)); while the epilogue warps drain accumulator i from
//     TMEM, the MMA warp already fills accumulator i^1 for the next tile;
//   * the epilogue goes through a padded shared-memory staging tile in 64-column chunks: residual / accumulate
//     operands are loaded into it with full 128-byte row segments, every thread then owns one row (TMEM lane) for the
//     arithmetic, and the fp16 results leave again as full row segments (4 rows per warp store).
// Same GemmParams / tensor maps / modes as gemm.cuh (which remains for fp32 outputs and tiny N).
#pragma once
#include "gemm.cuh"

namespace b200 {

template <int BLOCK_N>
struct Gemm2Cfg {
  static constexpr int A_BYTES = 16384;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int STG_LD = 72;                        // staging row stride in halves (144 B: conflict-free 16 B rows)
  static constexpr int STG_BYTES = 128 * STG_LD * 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 128 * 8 + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

// MODE is the epilogue (EPI_ROWMAJOR / EPI_GEGLU / EPI_HEADS) as a template parameter: each variant gets its own
// register allocation and instruction schedule (as one kernel with run-time mode tests, a change to the row-major
// epilogue moved the GEGLU variant by 10-15 %, profiles/r2/ops_profile*.txt)
template <int BLOCK_N, int MODE>
__global__ void __launch_bounds__(320, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmParams p, int m_tiles, int n_tiles) {
  using Cfg = Gemm2Cfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int LD = Cfg::STG_LD;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __half* stg = reinterpret_cast<__half*>(smem + STAGES * Cfg::STAGE_BYTES);
  long long* s_orow = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(stg) + Cfg::STG_BYTES);  // [128]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_orow + 128);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // 2
  uint64_t* tmem_empty = tmem_full + 2;       // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int total_tiles = m_tiles * n_tiles;
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
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tmem_full[s], 1);
        mbar_init(&tmem_empty[s], 8);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto tile_coords = [&](int tile, int& n0, int& x0, int& y0, int& b0) {
    const int n_tile = tile % n_tiles;
    const int m_tile = tile / n_tiles;
    n0 = n_tile * BLOCK_N;
    x0 = (m_tile % p.tiles_x) * p.tw;
    y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.th;
    b0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.tb;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n0, x0, y0, b0;
        tile_coords(tile, n0, x0, y0, b0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / cpb;
          const int cc = kb - tap * cpb;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const GemmTap t = p.taps[tap];
          tma_load_4d(sa, &tmA, &full_bar[stage], cc * 64, x0 + t.dx, y0 + t.dy, b0 + t.db);
          tma_load_3d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], cc * 64, t.wtap, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(128, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      mbar_wait(&tmem_empty[ab], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ab * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = make_desc_k_sw128(sa);
          const uint64_t bdesc = make_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          tc_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) tc_commit(&tmem_full[ab]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    // Two warps per TMEM lane quadrant: both own the same 32 rows, each takes 32 of the 64 columns of a chunk, so the
    // epilogue arithmetic (the bottleneck of the short-K GEMMs) is spread over all four SM sub-partitions twice.
    const int ew = warp - 2;                  // 0..7
    const int quad = warp & 3;
    const int half = ew >> 2;                 // column half inside a 64-column chunk
    const int r = quad * 32 + lane_id();      // row in tile == TMEM lane
    const int lane = lane_id();
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t stg_s = smem_u32(stg);
    const uint32_t my_s = stg_s + (r * LD + half * 32) * 2;
    __half* my = stg + r * LD + half * 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      int n0, x0, y0, b0;
      tile_coords(tile, n0, x0, y0, b0);
      const int n_tile = tile % n_tiles;
      const int xl = r % p.tw, yl = (r / p.tw) % p.th, bl = r / (p.tw * p.th);
      const int x = x0 + xl, y = y0 + yl, b = b0 + bl;
      const bool row_ok = (x < p.W) && (y < p.H) && (b < p.B);
      const long long orow = ((long long)b * p.OH + (y * p.sy + p.oy)) * p.OW + (x * p.sx + p.ox);
      asm volatile("bar.sync 1, 256;" ::: "memory");      // previous tile's staging reads are finished
      if (half == 0) s_orow[r] = row_ok ? orow : -1;
      // rows in the padding of a pixel brick (b >= B, x >= W, y >= H) never store; they take image 0 so that their
      // per-image channel-add reads stay inside the buffer
      const int img = (p.rows_per_img > 0 && row_ok) ? (int)(orow / p.rows_per_img) : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");      // s_orow visible
      // coalesced passes: warp ew moves rows ew*16 .. +15, 4 rows per instruction (8 lanes x 16 B per row)
      long long orws[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) orws[rr] = s_orow[ew * 16 + rr * 4 + (lane >> 3)];
      auto for_pieces = [&](auto&& fn) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) fn(ew * 16 + rr * 4 + (lane >> 3), lane & 7, orws[rr]);
      };
      // EPI_HEADS: (image, token) of the rows this thread touches, once per tile (32-bit: M < 2^31)
      int pim[4] = {0, 0, 0, 0}, ptok[4] = {0, 0, 0, 0}, tok_me = 0;
      if (MODE == EPI_HEADS) {
        const unsigned rpi = (unsigned)p.rows_per_img;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const unsigned o = orws[rr] >= 0 ? (unsigned)orws[rr] : 0u;
          pim[rr] = (int)(o / rpi);
          ptok[rr] = (int)(o - (unsigned)pim[rr] * rpi);
        }
        tok_me = (int)(orow - (long long)img * p.rows_per_img);
      }
      // residual / accumulate operand: fetched one chunk ahead into registers so its latency hides behind the
      // accumulator wait and the previous chunk's arithmetic
      const bool rd = (MODE == EPI_ROWMAJOR) && ((p.residual != nullptr) || p.accumulate_out);
      const __half* rsrc = p.residual ? p.residual : p.out;
      const int rlds = p.residual ? p.ldr : p.ldo;
      uint4 pre[4];
      auto load_res = [&](int nb) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int n = nb + (lane & 7) * 8;
          pre[rr] = (orws[rr] >= 0 && n + 8 <= p.N) ? *reinterpret_cast<const uint4*>(rsrc + orws[rr] * rlds + n)
                                                    : make_uint4(0, 0, 0, 0);
        }
      };
      if (rd) load_res(n0);
      mbar_wait(&tmem_full[ab], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * BLOCK_N + lane_off;

      if (MODE == EPI_GEGLU) {
        if constexpr (BLOCK_N == 128) {
          // 64 outputs per tile: value columns [0,64), gate columns [64,128); this thread: 32 of them
          uint32_t vv[32], gg[32];
          tmem_ld_x32(taddr + half * 32, vv);
          tmem_ld_x32(taddr + 64 + half * 32, gg);
          tmem_ld_wait();
          uint32_t o[16], pv[16], pg[16];
          // scalar __ldg bias reads (L1-resident, two per output): the 16-byte variant holds 64 more live registers
          // in front of the erf-GELU and measured 9-15 % slower on the three GEGLU shapes (profiles/r2/ops_profile.txt)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a0 = __uint_as_float(vv[2 * j]), a1 = __uint_as_float(vv[2 * j + 1]);
            float g0 = __uint_as_float(gg[2 * j]), g1 = __uint_as_float(gg[2 * j + 1]);
            if (p.bias) {
              a0 += __ldg(p.bias + n0 + half * 32 + 2 * j);
              a1 += __ldg(p.bias + n0 + half * 32 + 2 * j + 1);
              g0 += __ldg(p.bias + n0 + 64 + half * 32 + 2 * j);
              g1 += __ldg(p.bias + n0 + 64 + half * 32 + 2 * j + 1);
            }
            o[j] = pack_h2(a0 * gelu_erf(g0), a1 * gelu_erf(g1));
            pv[j] = pack_h2(a0, a1);
            pg[j] = pack_h2(g0, g1);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) sts128(my_s + q * 16, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
          if (p.pre && row_ok) {   // pre-activation dump (guidance forward only): direct, 64 B runs per row
            __half* pr = p.pre + orow * (long long)(2 * p.ldo) + n0 + half * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              *reinterpret_cast<uint4*>(pr + q * 8) = make_uint4(pv[4 * q], pv[4 * q + 1], pv[4 * q + 2], pv[4 * q + 3]);
              *reinterpret_cast<uint4*>(pr + 64 + q * 8) = make_uint4(pg[4 * q], pg[4 * q + 1], pg[4 * q + 2], pg[4 * q + 3]);
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          for_pieces([&](int row, int pc, long long orw) {
            if (orw >= 0)
              *reinterpret_cast<uint4*>(p.out + orw * p.ldo + n_tile * 64 + pc * 8) = lds128(stg_s + (row * LD + pc * 8) * 2);
          });
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 64) {
          if (n0 + c0 >= p.N) break;
          const int nb = n0 + c0;                          // first output column of this chunk
          if (c0 > 0) asm volatile("bar.sync 1, 256;" ::: "memory");   // staging free again
          if (rd) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
              sts128(stg_s + ((ew * 16 + rr * 4 + (lane >> 3)) * LD + (lane & 7) * 8) * 2, pre[rr]);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (c0 + 64 < BLOCK_N && nb + 64 < p.N) load_res(nb + 64);
          }
          uint32_t v[32];
          tmem_ld_x32(taddr + c0 + half * 32, v);
          tmem_ld_wait();
          const int nh = nb + half * 32;                   // first column of this thread's 32
          // per-column addend (bias + per-image channel add) of this thread's 32 columns: fetched once per chunk with
          // 16-byte loads instead of two predicated scalar loads per element (the epilogue of the short-K GEMMs is
          // bound by its instruction count)
          const bool has_add = (p.bias != nullptr) || (p.chan_add != nullptr);
          float add[32];
          if (has_add) {
            const float* cadd = p.chan_add ? p.chan_add + (long long)img * p.N : nullptr;
            if (nh + 32 <= p.N && (p.N & 3) == 0) {
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) {
                float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nh) + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (cadd) {
                  const float4 c4 = __ldg(reinterpret_cast<const float4*>(cadd + nh) + q4);
                  b4.x += c4.x; b4.y += c4.y; b4.z += c4.z; b4.w += c4.w;
                }
                add[4 * q4] = b4.x; add[4 * q4 + 1] = b4.y; add[4 * q4 + 2] = b4.z; add[4 * q4 + 3] = b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int nn = nh + j;
                float a = 0.f;
                if (nn < p.N) {
                  if (p.bias) a += __ldg(p.bias + nn);
                  if (cadd) a += __ldg(cadd + nn);
                }
                add[j] = a;
              }
            }
          }
          const bool scaled = p.alpha != 1.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a = __uint_as_float(v[g * 8 + j]);
              if (scaled) a *= p.alpha;
              if (has_add) a += add[g * 8 + j];
              f[j] = a;
            }
            if (rd) {
              uint4 rr = lds128(my_s + g * 16);
              const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 t2 = __half22float2(rh[j]);
                f[2 * j] += t2.x;
                f[2 * j + 1] += t2.y;
              }
            }
            sts128(my_s + g * 16,
                   make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7])));
          }
          if (MODE == EPI_HEADS) {
            // transposed slabs: this thread's row is one token, lanes of a warp are consecutive tokens -> coalesced.
            // (projection, head, column-in-head) of the first 8-column group, then stepped without divisions
            if (row_ok) {
              int which = nh / p.C;
              const int cc = nh - which * p.C;
              int head = cc / p.d, j0 = cc - head * p.d;
              which += p.which0;
#pragma unroll 1
              for (int g = 0; g < 4; ++g) {
                if (nh + g * 8 >= p.N) break;
                if (p.tr[which]) {
                  const long long ta = p.tr_alloc[which];
                  __half* dst = p.tr[which] + (((long long)img * p.heads + head) * p.d16 + j0) * ta + tok_me;
                  const uint4 pk = lds128(my_s + g * 16);
                  const __half* hv = reinterpret_cast<const __half*>(&pk);
#pragma unroll
                  for (int j = 0; j < 8; ++j) dst[j * ta] = hv[j];
                }
                j0 += 8;
                if (j0 >= p.d) {
                  j0 = 0;
                  if (++head == p.heads) { head = 0; ++which; }
                }
              }
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (MODE == EPI_ROWMAJOR) {
            const bool vec_ok = (p.ldo % 8 == 0);
            for_pieces([&](int row, int pc, long long orw) {
              const int n = nb + pc * 8;
              if (orw < 0 || n >= p.N) return;
              if (vec_ok && n + 8 <= p.N) {
                *reinterpret_cast<uint4*>(p.out + orw * p.ldo + n) = lds128(stg_s + (row * LD + pc * 8) * 2);
              } else {
                const __half* s = stg + row * LD + pc * 8;
                for (int j = 0; j < 8 && n + j < p.N; ++j) p.out[orw * p.ldo + n + j] = s[j];
              }
            });
          } else {  // EPI_HEADS row-major slabs: 16-byte pieces never straddle a head (d % 8 == 0)
            const int n = nb + (lane & 7) * 8;            // the same column group for this thread's four rows
            if (n < p.N) {
              const int wh = n / p.C;
              const int cc = n - wh * p.C;
              const int head = cc / p.d, j0 = cc - head * p.d;
              const int which = p.which0 + wh;
              if (p.rm[which]) {
                const long long ra = p.rm_alloc[which];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                  if (orws[rr] < 0) continue;
                  const int row = ew * 16 + rr * 4 + (lane >> 3);
                  *reinterpret_cast<uint4*>(p.rm[which] + (((long long)pim[rr] * p.heads + head) * ra + ptok[rr]) * (long long)p.dp + j0) =
                      lds128(stg_s + (row * LD + (lane & 7) * 8) * 2);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[ab]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
