// Self-attention (and un-guided cross-attention) forward on tcgen05: S = Q K^T and O = P V as 128-row tcgen05 tiles
// with S/O accumulators in TMEM, softmax by one thread per query row straight out of TMEM.
//
// Two-pass softmax: pass 1 streams K and reduces the exact row max / row sum (no P, no V traffic); pass 2 recomputes S,
// writes normalised fp16 P to shared memory (K-major SW128, the A operand of P.V) and accumulates O in TMEM with no
// rescaling.  The row statistic L2 = m + log2(l) (log2 domain) is stored for the backward kernels.
//
// Operand slabs (written by the projection GEMM's EPI_HEADS epilogue):
//   Q, K : [B*heads, n_alloc, dp]   row-major, dp = head_dim rounded up to 64 (zero padded)
//   V^T  : [B*heads, d16, nk_alloc] d-major,   d16 = head_dim rounded up to 16
// Warp roles: warp 0 TMA, warp 1 MMA issue (+TMEM alloc), warps 2..5 softmax / epilogue (thread == query row).
#pragma once
#include "ptx.cuh"

namespace b200 {

struct AttnParams {
  int heads, nq, nk;       // valid lengths
  int nq_alloc, nk_alloc;  // slab rows
  int d;                   // true head dim
  float scale_log2;        // softmax scale * log2(e)
  __half* out;             // [B*nq, ldo], head h at columns [h*d, h*d+d)
  int ldo;
  float* lse2;             // [B*heads, nq_alloc] or null
};

template <int DPB, int D16, int STAGES>
struct AttnCfg {
  static constexpr int Q_BYTES = DPB * 16384;
  static constexpr int K_BYTES = DPB * 16384;
  static constexpr int V_ATOM = D16 * 128;
  static constexpr int V_BYTES = 2 * V_ATOM;
  static constexpr int P_BYTES = 2 * 16384;
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * (K_BYTES + V_BYTES) + P_BYTES + 1024 + 256;
};

template <int DPB, int D16, int STAGES>
__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<DPB, D16, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + STAGES * Cfg::K_BYTES;
  uint8_t* sP = sV + STAGES * Cfg::V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // STAGES
  uint64_t* kv_empty = kv_full + STAGES;   // STAGES
  uint64_t* s_full = kv_empty + STAGES;    // 2
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* p_full = s_empty + 2;          // 1
  uint64_t* p_empty = p_full + 1;          // 1
  uint64_t* o_full = p_empty + 1;          // 1
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x;
  const int bh = blockIdx.y;
  const int nkv = (p.nk + 127) >> 7;
  const int total = 2 * nkv;

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
      for (int s = 0; s < 2; ++s) {
        mbar_init(&s_full[s], 1);
        mbar_init(&s_empty[s], 4);
      }
      mbar_init(p_full, 4);
      mbar_init(p_empty, 1);
      mbar_init(o_full, 1);
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
  const uint32_t tO = tmem_base + 256;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
      for (int a = 0; a < DPB; ++a) tma_load_3d(sQ + a * 16384, &tmQ, q_full, a * 64, qt * 128, bh);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total; ++it) {
        const int j = it % nkv;
        const int pass = it / nkv;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::K_BYTES + (pass ? Cfg::V_BYTES : 0));
        for (int a = 0; a < DPB; ++a)
          tma_load_3d(sK + stage * Cfg::K_BYTES + a * 16384, &tmK, &kv_full[stage], a * 64, j * 128, bh);
        if (pass) {
          for (int a = 0; a < 2; ++a)
            tma_load_3d(sV + stage * Cfg::V_BYTES + a * Cfg::V_ATOM, &tmVt, &kv_full[stage], j * 128 + a * 64, 0, bh);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_f16(128, D16);
    mbar_wait(q_full, 0);
    const uint32_t qaddr = smem_u32(sQ);
    auto issue_S = [&](int it) {
      const int stage = it % STAGES;
      const uint32_t kphase = (it / STAGES) & 1;
      const int sb = it & 1;
      const uint32_t sphase = (it >> 1) & 1;
      mbar_wait(&kv_full[stage], kphase);
      mbar_wait(&s_empty[sb], sphase ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t kaddr = smem_u32(sK + stage * Cfg::K_BYTES);
#pragma unroll
        for (int a = 0; a < DPB; ++a) {
          const uint64_t ad = make_desc_k_sw128(qaddr + a * 16384);
          const uint64_t bd = make_desc_k_sw128(kaddr + a * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tS[sb], ad + k * 2, bd + k * 2, idesc_s, (a | k) ? 1u : 0u);
        }
        tc_commit(&s_full[sb]);
        if (it < nkv) tc_commit(&kv_empty[stage]);  // pass 1: K tile is free once S is computed
      }
      __syncwarp();
    };
    auto issue_PV = [&](int it) {
      const int j = it - nkv;
      const int stage = it % STAGES;
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t paddr = smem_u32(sP);
        const uint32_t vaddr = smem_u32(sV + stage * Cfg::V_BYTES);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const uint64_t ad = make_desc_k_sw128(paddr + a * 16384);
          const uint64_t bd = make_desc_k_sw128(vaddr + a * Cfg::V_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tO, ad + k * 2, bd + k * 2, idesc_o, (j | a | k) ? 1u : 0u);
        }
        tc_commit(&kv_empty[stage]);
        tc_commit(p_empty);
        if (it == total - 1) tc_commit(o_full);
      }
      __syncwarp();
    };
    if constexpr (STAGES >= 2) {
      issue_S(0);
      for (int it = 0; it < total; ++it) {
        if (it + 1 < total) issue_S(it + 1);
        if (it >= nkv) issue_PV(it);
      }
    } else {
      for (int it = 0; it < total; ++it) {
        issue_S(it);
        if (it >= nkv) issue_PV(it);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane_id();
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m = -INFINITY, l = 0.f;
    // ---------------- pass 1: exact row max and row sum
    for (int it = 0; it < nkv; ++it) {
      const int sb = it & 1;
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      tc_fence_after();
      const int kbase = it * 128;
      float tmax = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tS[sb] + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kbase + c0 + i < p.nk) tmax = fmaxf(tmax, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m, tmax * p.scale_log2);
      float acc = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tS[sb] + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kbase + c0 + i < p.nk) acc += exp2f(__uint_as_float(v[i]) * p.scale_log2 - m_new);
      }
      l = l * exp2f(m - m_new) + acc;
      m = m_new;
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&s_empty[sb]);
    }
    const float inv_l = 1.f / l;
    // ---------------- pass 2: P = 2^(s' - m) / l  -> smem (fp16, SW128 K-major), O += P V on the tensor core
    for (int j = 0; j < nkv; ++j) {
      const int it = nkv + j;
      const int sb = it & 1;
      mbar_wait(&s_full[sb], (it >> 1) & 1);
      mbar_wait(p_empty, (j & 1) ^ 1);
      tc_fence_after();
      const int kbase = j * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tS[sb] + lane_off + c0, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k0 = kbase + c0 + 2 * i;
          const float a = (k0 < p.nk) ? exp2f(__uint_as_float(v[2 * i]) * p.scale_log2 - m) * inv_l : 0.f;
          const float b = (k0 + 1 < p.nk) ? exp2f(__uint_as_float(v[2 * i + 1]) * p.scale_log2 - m) * inv_l : 0.f;
          pk[i] = pack_h2(a, b);
        }
        uint8_t* atom = sP + (c0 >> 6) * 16384;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 st = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0 + q)) = st;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) {
        mbar_arrive(&s_empty[sb]);
        mbar_arrive(p_full);
      }
    }
    // ---------------- epilogue
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qrow = qt * 128 + r;
    const int b = bh / p.heads, h = bh % p.heads;
    const bool ok = qrow < p.nq;
    __half* orow = p.out + ((long long)b * p.nq + qrow) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c0 = 0; c0 < D16; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tO + lane_off + c0, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c0 + g * 8 < p.d) {
            uint4 st;
            st.x = pack_h2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            st.y = pack_h2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            st.z = pack_h2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            st.w = pack_h2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + c0 + g * 8) = st;
          }
        }
      }
    }
    if (ok && p.lse2) p.lse2[(long long)bh * p.nq_alloc + qrow] = m + log2f(l);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
