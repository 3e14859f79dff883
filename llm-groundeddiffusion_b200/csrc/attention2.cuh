// Self-attention forward, second generation (head_dim <= 64): two 128-row query tiles per CTA in ping-pong, single-pass
// online softmax with lazy rescaling.
//
// Why (profiles/r1_launches_step_summary.md): the first kernel spent ~5.5 k cycles per 128x128 score tile per pass - one
// softmax warp per SM sub-partition (no latency hiding) and two exponentials per score.  Here eight softmax warps
// (two per sub-partition) each own one query row per thread of tile A or B; while tile A's rows are in the softmax, the
// tensor core works on tile B (S_B = Q_B K^T, O_B += P_B V) and vice versa.  Every score is exponentiated once:
//     m_ref is only raised when the tile maximum exceeds it by more than 8 (log2 units), in which case the O accumulator
//     in TMEM and the running sum are rescaled by 2^(m_old - m_new) (tcgen05.ld / .st of the row); otherwise P stays
//     relative to the stale reference (<= 2^8, safe in fp16) and the final 1/l normalises exactly.
// K/V tiles are shared by both query tiles (half the L2->smem traffic per row).  The score row of a tile is read from
// TMEM once into registers (TMEM reads run at 64 B/clk/SM and are what bounds this kernel at head_dim 40: 64 KB per
// 128x128 tile), after which S(j+1) is issued into the same TMEM columns while the exponentials of S(j) run.  Same
// slab layout and outputs as attn_fwd_kernel (attention.cuh).
#pragma once
#include "attention.cuh"

namespace b200 {

// P lives in tensor memory (written by tcgen05.st, consumed as the A operand of P.V): the P tile never touches the
// shared-memory port (32 KB written + 32 KB read per 128x128 tile otherwise), which the S = Q.K^T operand reads and the
// TMA writes need, and the 64 KB it would occupy go to deeper K/V staging.
template <int D16, int STAGES>
struct Attn2Cfg {
  static constexpr int Q_BYTES = 2 * 16384;             // two query tiles, dp = 64
  static constexpr int K_BYTES = 16384;
  static constexpr int V_ATOM = D16 * 128;
  static constexpr int V_BYTES = 2 * V_ATOM;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * (K_BYTES + V_BYTES) + 1024 + 256;
  static constexpr int O_STRIDE = (D16 + 63) / 64 * 64;
};

template <int D16, int STAGES>
__global__ void __launch_bounds__(320, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ AttnParams p) {
  using Cfg = Attn2Cfg<D16, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + STAGES * Cfg::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * Cfg::V_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // STAGES
  uint64_t* kv_empty = kv_full + STAGES;   // STAGES
  uint64_t* s_full = kv_empty + STAGES;    // 2 (per query tile)
  uint64_t* p_full = s_full + 2;           // 2
  uint64_t* pv_done = p_full + 2;          // 2
  uint64_t* s_free = pv_done + 2;          // 2: the score row of step j sits in registers, S(j+1) may overwrite TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.y;
  const int nkv = (p.nk + 127) >> 7;
  const int ntile = (p.nq - q0 > 128) ? 2 : 1;   // second tile entirely out of range -> skip it

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmVt);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(&s_full[t], 1);
        mbar_init(&p_full[t], 4);
        mbar_init(&pv_done[t], 1);
        mbar_init(&s_free[t], 4);
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
  const uint32_t tS[2] = {tmem_base, tmem_base + 128};
  const uint32_t tO[2] = {tmem_base + 256, tmem_base + 256 + Cfg::O_STRIDE};
  const uint32_t tP[2] = {tmem_base + 384, tmem_base + 448};   // 128 keys = 64 packed fp16x2 columns per query tile

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ntile * 16384);
      for (int t = 0; t < ntile; ++t) tma_load_3d(sQ + t * 16384, &tmQ, q_full, 0, q0 + t * 128, bh);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::K_BYTES + Cfg::V_BYTES);
        tma_load_3d(sK + stage * Cfg::K_BYTES, &tmK, &kv_full[stage], 0, j * 128, bh);
        for (int a = 0; a < 2; ++a)
          tma_load_3d(sV + stage * Cfg::V_BYTES + a * Cfg::V_ATOM, &tmVt, &kv_full[stage], j * 128 + a * 64, 0, bh);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_f16(128, D16);
    mbar_wait(q_full, 0);
    auto issue_S = [&](int t, int j) {
      const int stage = j % STAGES;
      if (elect_one()) {
        const uint64_t ad = make_desc_k_sw128(smem_u32(sQ) + t * 16384);
        const uint64_t bd = make_desc_k_sw128(smem_u32(sK) + stage * Cfg::K_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tS[t], ad + k * 2, bd + k * 2, idesc_s, k ? 1u : 0u);
        tc_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_PV = [&](int t, int j) {
      const int stage = j % STAGES;
      if (elect_one()) {
        const uint32_t vaddr = smem_u32(sV) + stage * Cfg::V_BYTES;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const uint64_t bd = make_desc_k_sw128(vaddr + a * Cfg::V_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k)      // 16 keys per MMA: 8 packed columns of P, 32 bytes of the V^T atom
            umma_f16_ts(tO[t], tP[t] + (a * 4 + k) * 8, bd + k * 2, idesc_o, (j | a | k) ? 1u : 0u);
        }
        tc_commit(&pv_done[t]);
      }
      __syncwarp();
    };
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    for (int t = 0; t < ntile; ++t) issue_S(t, 0);
    for (int j = 0; j < nkv; ++j) {
      const int stage = j % STAGES;
      // S(j+1) is issued as soon as the softmax warps hold S(j) in registers: it runs under their exponentials, so
      // the next score tile is ready the moment they finish the current one
      if (j + 1 < nkv) {
        mbar_wait(&kv_full[(j + 1) % STAGES], ((j + 1) / STAGES) & 1);
        for (int t = 0; t < ntile; ++t) {
          mbar_wait(&s_free[t], j & 1);
          tc_fence_after();
          issue_S(t, j + 1);
        }
      }
      for (int t = 0; t < ntile; ++t) {
        mbar_wait(&p_full[t], j & 1);
        tc_fence_after();
        issue_PV(t, j);
      }
      if (elect_one()) tc_commit(&kv_empty[stage]);   // every MMA that read stage j has been issued before this
      __syncwarp();
    }
  } else {
    const int t = (warp - 2) >> 2;          // query tile of this warp
    const int quad = warp & 3;
    const int r = quad * 32 + lane_id();
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    if (t < ntile) {
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        const int kbase = j * 128;
        const bool full = (kbase + 128 <= p.nk);
        // ---- the whole score row of this tile in registers: one TMEM read (four loads in flight, one wait)
        uint32_t v[128];
        tmem_ld_x32(tS[t] + lane_off, v);
        tmem_ld_x32(tS[t] + lane_off + 32, v + 32);
        tmem_ld_x32(tS[t] + lane_off + 64, v + 64);
        tmem_ld_x32(tS[t] + lane_off + 96, v + 96);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&s_free[t]);
        // ---- tile maximum (4 independent chains)
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        if (full) {
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            mx0 = max3(mx0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            mx1 = max3(mx1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            mx2 = max3(mx2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            mx3 = max3(mx3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (kbase + i < p.nk) mx0 = fmaxf(mx0, __uint_as_float(v[i]));
        }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
        const bool need = m_tile > m_ref + 8.f;
        const float m_new = need ? m_tile : m_ref;
        // ---- P(j-1) consumed and O stable before we touch either
        if (j > 0) {
          mbar_wait(&pv_done[t], (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx(m_ref - m_new) : 1.f;
#pragma unroll 1
            for (int c0 = 0; c0 < D16; c0 += 16) {
              uint32_t o[16];
              tmem_ld_x16(tO[t] + lane_off + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(tO[t] + lane_off + c0, o);
            }
            tmem_st_wait();
            l *= alpha;
          }
        }
        m_ref = m_new;
        // ---- P = 2^(s' - m_ref) -> fp16 smem, row sum
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float e0 = ex2_approx(fmaf(__uint_as_float(v[c0 + 2 * i]), p.scale_log2, -m_ref));
            float e1 = ex2_approx(fmaf(__uint_as_float(v[c0 + 2 * i + 1]), p.scale_log2, -m_ref));
            if (!full) {
              if (kbase + c0 + 2 * i >= p.nk) e0 = 0.f;
              if (kbase + c0 + 2 * i + 1 >= p.nk) e1 = 0.f;
            }
            a0 += e0;
            a1 += e1;
            pk[i] = pack_h2(e0, e1);
          }
          tmem_st_x16(tP[t] + lane_off + (c0 >> 1), pk);   // keys c0..c0+31 -> 16 packed columns
        }
        tmem_st_wait();
        l += a0 + a1;
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&p_full[t]);
      }
      // ---------------- epilogue: O / l
      mbar_wait(&pv_done[t], (nkv - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.f / l;
      const int qrow = q0 + t * 128 + r;
      const int b = bh / p.heads, h = bh % p.heads;
      const bool ok = qrow < p.nq;
      __half* orow = p.out + ((long long)b * p.nq + qrow) * p.ldo + h * p.d;
#pragma unroll 1
      for (int c0 = 0; c0 < D16; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(tO[t] + lane_off + c0, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g)
            if (c0 + g * 8 < p.d) {
              uint4 st;
              st.x = pack_h2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
              st.y = pack_h2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
              st.z = pack_h2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
              st.w = pack_h2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + c0 + g * 8) = st;
            }
        }
      }
      if (ok && p.lse2) p.lse2[(long long)bh * p.nq_alloc + qrow] = m_ref + log2f(l);
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int D16>
inline void launch_attn_fwd2_t(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVt,
                               const AttnParams& p, int nq, int BH, cudaStream_t st) {
  constexpr int ST = 4;
  using Cfg = Attn2Cfg<D16, ST>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "smem budget");
  static_assert(256 + 2 * Cfg::O_STRIDE + 128 <= 512, "TMEM budget");
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(attn_fwd2_kernel<D16, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    done = true;
  }
  dim3 grid((nq + 255) / 256, BH, 1);
  attn_fwd2_kernel<D16, ST><<<grid, 320, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmVt, p);
}

}  // namespace b200
