// Implicit GEMM, third generation (row-major fp16 outputs): persistent, double-buffered TMEM, and an epilogue that
// touches global memory only through TMA.
//
// Why (profiles/: ncu of gemm2 on M=65536,N=320,K=320 shows no hot instruction any more, 30 % issue activity and 10 %
// tensor activity: the short-K GEMMs are bound by the LATENCY chain of the epilogue - residual loads, named barriers,
// staged copies - not by any throughput).  Here
//   * the producer warp TMA-loads the residual (or accumulate) tile of the NEXT tile into a swizzled shared buffer
//     while the current tile is still in the tensor core, so the epilogue never waits on HBM;
//   * epilogue threads (two warps per TMEM lane quadrant, 32 columns each) read their accumulator row, add bias /
//     time-embedding / residual from the buffer, write fp16 back INTO the same buffer;
//   * one thread issues TMA stores of the 128x64 atoms (async, coalesced, clipped at the tensor edge) and releases the
//     buffer once the bulk group has been read.
// Same mainloop, tensor maps and GemmParams as gemm2.cuh.  Requires the identity output-pixel mapping (sy = sx = 1).
#pragma once
#include "gemm.cuh"

namespace b200 {

template <int BLOCK_N>
struct Gemm3Cfg {
  static constexpr int A_BYTES = 16384;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 3 : 4;
  static constexpr int ATOMS = BLOCK_N / 64;
  static constexpr int EB_BYTES = ATOMS * 16384;            // one epilogue buffer: ATOMS x (128 rows x 128 B)
  static constexpr int NEB = (BLOCK_N == 256) ? 1 : 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NEB * EB_BYTES + 1024 + 256;
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(320, 1)
gemm3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ GemmParams p, int m_tiles, int n_tiles) {
  using Cfg = Gemm3Cfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NEB = Cfg::NEB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* eb_base = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(eb_base + NEB * Cfg::EB_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // 2
  uint64_t* tmem_empty = tmem_full + 2;       // 2 (count 8)
  uint64_t* res_full = tmem_empty + 2;        // NEB
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + NEB);

  const int warp = threadIdx.x >> 5;
  const int total_tiles = m_tiles * n_tiles;
  const int cpb = (p.Cin + 63) >> 6;
  const int num_kb = cpb * p.ntaps;
  const bool has_res = (p.residual != nullptr) || p.accumulate_out;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    prefetch_tmap(&tmR);
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
      for (int s = 0; s < NEB; ++s) mbar_init(&res_full[s], 1);
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
  auto atoms_of = [&](int n0) {
    const int rem = p.N - n0;
    const int a = (rem + 63) >> 6;
    return a < Cfg::ATOMS ? a : Cfg::ATOMS;
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
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int lane = lane_id();
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool leader = (ew == 0) && (lane == 0);
    // the leader also feeds the residual tiles: buffer (it+1) % NEB is free as soon as the store that last read it has
    // drained, which the leader itself observes (bulk async-groups are per thread)
    auto load_residual = [&](int tile, int buf) {
      int n0, x0, y0, b0;
      tile_coords(tile, n0, x0, y0, b0);
      const int na = atoms_of(n0);
      mbar_arrive_expect_tx(&res_full[buf], na * 16384);
      for (int a = 0; a < na; ++a)
        tma_load_4d(eb_base + buf * Cfg::EB_BYTES + a * 16384, &tmR, &res_full[buf], n0 + a * 64, x0, y0, b0);
    };
    if (leader && has_res && (int)blockIdx.x < total_tiles) load_residual(blockIdx.x, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const int eb = it % NEB;
      const int next = tile + gridDim.x;
      int n0, x0, y0, b0;
      tile_coords(tile, n0, x0, y0, b0);
      const int na = atoms_of(n0);
      const uint32_t ebuf = smem_u32(eb_base + eb * Cfg::EB_BYTES);
      if (leader && it > 0) tma_store_wait_read<0>();    // every earlier store has been read out of shared memory
      if constexpr (NEB == 2) {
        if (leader && has_res && next < total_tiles) load_residual(next, (it + 1) % NEB);
      }
      if (!has_res && it > 0) asm volatile("bar.sync 1, 256;" ::: "memory");   // buffer free (leader waited above)
      int img = 0;
      if (p.chan_add) {
        const int xl = r % p.tw, yl = (r / p.tw) % p.th, bl = r / (p.tw * p.th);
        const long long orow = ((long long)(b0 + bl) * p.OH + (y0 + yl)) * p.OW + (x0 + xl);
        img = (b0 + bl < p.B) ? (int)(orow / p.rows_per_img) : 0;
      }
      if (has_res) mbar_wait(&res_full[eb], (it / NEB) & 1);   // residual tile landed (TMA -> smem)
      mbar_wait(&tmem_full[ab], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * BLOCK_N + lane_off;
#pragma unroll 1
      for (int a = 0; a < na; ++a) {
        uint32_t v[32];
        tmem_ld_x32(taddr + a * 64 + half * 32, v);
        tmem_ld_wait();
        const int nh = n0 + a * 64 + half * 32;
        const uint32_t arow = ebuf + a * 16384;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float acc = __uint_as_float(v[g * 8 + j]) * p.alpha;
            const int nn = nh + g * 8 + j;
            if (nn < p.N) {
              if (p.bias) acc += __ldg(p.bias + nn);
              if (p.chan_add) acc += __ldg(p.chan_add + (long long)img * p.N + nn);
            }
            f[j] = acc;
          }
          const uint32_t paddr = arow + sw128_offset(r, half * 4 + g);
          if (has_res) {
            const uint4 rr = lds128(paddr);
            const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 t2 = __half22float2(rh[j]);
              f[2 * j] += t2.x;
              f[2 * j + 1] += t2.y;
            }
          }
          sts128(paddr, make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7])));
        }
      }
      fence_proxy_async();                 // generic-proxy writes -> visible to the TMA store
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[ab]);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the whole tile is in the buffer
      if (leader) {
        for (int a = 0; a < na; ++a)
          tma_store_4d(&tmC, eb_base + eb * Cfg::EB_BYTES + a * 16384, n0 + a * 64, x0, y0, b0);
        tma_store_commit();
        if constexpr (NEB == 1) {
          if (has_res && next < total_tiles) {
            tma_store_wait_read<0>();
            load_residual(next, 0);
          }
        }
      }
    }
    if (leader) tma_store_wait_all<0>();   // all output bytes are globally visible before the kernel ends
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
