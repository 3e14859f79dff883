// Attention backward on tcgen05 (dgrad only: weights are frozen, so only dQ / dK / dV w.r.t. activations exist).
//
//   attn_bwd_dq_kernel  : per query tile, loop over KV tiles:  S = Q K^T, dP = dO V^T  (TMEM) ->
//                         dS = P * (dP + dP_extra - delta) * scale  (registers -> fp16 smem) ->  dQ += dS K  (TMEM)
//   attn_bwd_dkv_kernel : per KV tile, loop over query tiles:  S^T = K Q^T, dP^T = V dO^T ->
//                         P^T, dS^T (fp16 smem) ->  dV += P^T dO,  dK += dS^T Q
//   attn_delta_kernel   : delta[bh, q] = sum_j dO[q, j] * O[q, j]
// P is recomputed from the stored log2-domain row statistic L2 (no n x n tensor ever touches HBM - the reference's
// math-SDPA guidance pass materialises it, models/pipelines.py:167).  `dP_extra` carries d(loss)/dP of the guidance
// loss for cross-attention layers; with a single KV tile delta is formed in-kernel.
// Slab layouts as in attention.cuh, plus row-major V / dO and transposed K^T / Q^T / dO^T slabs.
#pragma once
#include "ptx.cuh"

namespace b200 {

struct AttnBwdParams {
  int heads, nq, nk;
  int nq_alloc, nk_alloc;
  int d;
  float scale, scale_log2;
  const float* lse2;    // [BH, nq_alloc]
  const float* delta;   // [BH, nq_alloc] or null (then computed in-kernel; requires one KV tile)
  const float* extra;   // [BH, nq, ext_ld] fp32 additive dP term or null
  int ext_ld;
  int has_dO;
  __half* dq; int ld_dq;   // [B*nq, ld] head h at columns h*d
  __half* dk; int ld_dk;   // [B*nk_store, ld]
  __half* dv; int ld_dv;
  int nk_store;            // rows of dk/dv to write (<= nk); row stride of the output batch is nk_store
};

template <int DPB, int D16, int KVT, int STAGES>
struct AttnDqCfg {
  static constexpr int Q_BYTES = DPB * 16384;
  static constexpr int K_ATOM = KVT * 128;
  static constexpr int K_BYTES = DPB * K_ATOM;
  static constexpr int KT_ATOM = D16 * 128;
  static constexpr int KT_BYTES = (KVT / 64) * KT_ATOM;
  static constexpr int STAGE_BYTES = 2 * K_BYTES + KT_BYTES;
  static constexpr int DS_BYTES = (KVT / 64) * 16384;
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + STAGES * STAGE_BYTES + DS_BYTES + 1024 + 256;
};

template <int DPB, int D16, int KVT, int STAGES>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmKt, const __grid_constant__ AttnBwdParams p) {
  using Cfg = AttnDqCfg<DPB, D16, KVT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + Cfg::Q_BYTES;
  uint8_t* sStage = sdO + Cfg::Q_BYTES;
  uint8_t* sdS = sStage + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + Cfg::DS_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + STAGES;
  uint64_t* sdp_full = kv_empty + STAGES;
  uint64_t* sdp_empty = sdp_full + 1;
  uint64_t* ds_full = sdp_empty + 1;
  uint64_t* ds_empty = ds_full + 1;
  uint64_t* dq_full = ds_empty + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dq_full + 1);

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x;
  const int bh = blockIdx.y;
  const int nkv = (p.nk + KVT - 1) / KVT;
  // without a precomputed delta the KV tiles are walked twice: pass A reduces delta = sum_k P*dP, pass B forms dS
  const int first_b = p.delta ? 0 : nkv;
  const int total = first_b + nkv;

  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 1);
      }
      mbar_init(sdp_full, 1);
      mbar_init(sdp_empty, 8);   // eight softmax warps: two per TMEM lane quadrant, each takes half the key columns
      mbar_init(ds_full, 8);
      mbar_init(ds_empty, 1);
      mbar_init(dq_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdQ = tmem_base + 256;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES * (p.has_dO ? 2 : 1));
      for (int a = 0; a < DPB; ++a) {
        tma_load_3d(sQ + a * 16384, &tmQ, q_full, a * 64, qt * 128, bh);
        if (p.has_dO) tma_load_3d(sdO + a * 16384, &tmdO, q_full, a * 64, qt * 128, bh);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < total; ++it) {
        const int j = it % nkv;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sK = sStage + stage * Cfg::STAGE_BYTES;
        uint8_t* sV = sK + Cfg::K_BYTES;
        uint8_t* sKt = sV + Cfg::K_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], Cfg::K_BYTES * (p.has_dO ? 2 : 1) + Cfg::KT_BYTES);
        for (int a = 0; a < DPB; ++a) {
          tma_load_3d(sK + a * Cfg::K_ATOM, &tmK, &kv_full[stage], a * 64, j * KVT, bh);
          if (p.has_dO) tma_load_3d(sV + a * Cfg::K_ATOM, &tmV, &kv_full[stage], a * 64, j * KVT, bh);
        }
        for (int a = 0; a < KVT / 64; ++a)
          tma_load_3d(sKt + a * Cfg::KT_ATOM, &tmKt, &kv_full[stage], j * KVT + a * 64, 0, bh);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, KVT);
    constexpr uint32_t idesc_q = make_idesc_f16(128, D16);
    mbar_wait(q_full, 0);
    int stage = 0;
    uint32_t phase = 0;
    // S / dP of tile it+1 are issued as soon as the softmax warps hold tile it in registers (sdp_empty), i.e. they run
    // under the element-wise work of tile it; dQ(it) follows when dS(it) is in shared memory.
    auto issue_sdp = [&](int it, int stage) {
      const bool pass_b = it >= first_b;
      const uint32_t kaddr = smem_u32(sStage + stage * Cfg::STAGE_BYTES);
      const uint32_t vaddr = kaddr + Cfg::K_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < DPB; ++a) {
          const uint64_t qd = make_desc_k_sw128(smem_u32(sQ) + a * 16384);
          const uint64_t kd = make_desc_k_sw128(kaddr + a * Cfg::K_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tS, qd + k * 2, kd + k * 2, idesc_s, (a | k) ? 1u : 0u);
        }
        if (p.has_dO) {
#pragma unroll
          for (int a = 0; a < DPB; ++a) {
            const uint64_t od = make_desc_k_sw128(smem_u32(sdO) + a * 16384);
            const uint64_t vd = make_desc_k_sw128(vaddr + a * Cfg::K_ATOM);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tdP, od + k * 2, vd + k * 2, idesc_s, (a | k) ? 1u : 0u);
          }
        }
        tc_commit(sdp_full);
        if (!pass_b) tc_commit(&kv_empty[stage]);  // pass A: the stage is free once S / dP exist
      }
      __syncwarp();
    };
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    issue_sdp(0, 0);
    for (int it = 0; it < total; ++it) {
      const bool pass_b = it >= first_b;
      const int j = it - first_b;  // tile index inside pass B
      const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
      const uint32_t nphase = (nstage == 0) ? (phase ^ 1) : phase;
      auto next_sdp = [&]() {
        if (it + 1 < total) {
          mbar_wait(&kv_full[nstage], nphase);
          mbar_wait(sdp_empty, it & 1);
          tc_fence_after();
          issue_sdp(it + 1, nstage);
        }
      };
      if constexpr (STAGES >= 2) next_sdp();   // with a single stage the next tile only lands after dQ(it) frees it
      const uint32_t ktaddr = smem_u32(sStage + stage * Cfg::STAGE_BYTES) + 2 * Cfg::K_BYTES;
      if (pass_b) {
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int a = 0; a < KVT / 64; ++a) {
            const uint64_t sd = make_desc_k_sw128(smem_u32(sdS) + a * 16384);
            const uint64_t td = make_desc_k_sw128(ktaddr + a * Cfg::KT_ATOM);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tdQ, sd + k * 2, td + k * 2, idesc_q, (j | a | k) ? 1u : 0u);
          }
          tc_commit(&kv_empty[stage]);
          tc_commit(ds_empty);
          if (j == nkv - 1) tc_commit(dq_full);
        }
        __syncwarp();
      }
      if constexpr (STAGES < 2) next_sdp();
      stage = nstage;
      phase = nphase;
    }
  } else {
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;          // column half handled by this warp in pass B
    const int r = quad * 32 + lane_id();
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int qrow = qt * 128 + r;
    const bool ok = qrow < p.nq;
    const float L2 = ok ? p.lse2[(long long)bh * p.nq_alloc + qrow] : 0.f;
    const float* ext = (p.extra && ok) ? p.extra + ((long long)bh * p.nq + qrow) * p.ext_ld : nullptr;
    float delta = (p.delta && ok) ? p.delta[(long long)bh * p.nq_alloc + qrow] : 0.f;
    for (int it = 0; it < total; ++it) {
      const bool pass_b = it >= first_b;
      const int j = it - first_b;
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      const int kbase = (it % nkv) * KVT;
      if (!pass_b) {
        // pass A: delta += sum_k P * dP_total over this tile
#pragma unroll 1
        for (int c0 = 0; c0 < KVT; c0 += 32) {
          uint32_t s[32], g[32];
          tmem_ld_x32(tS + lane_off + c0, s);
          if (p.has_dO) tmem_ld_x32(tdP + lane_off + c0, g);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kk = kbase + c0 + i;
            if (ok && kk < p.nk) {
              const float pr = exp2f(__uint_as_float(s[i]) * p.scale_log2 - L2);
              float dp = p.has_dO ? __uint_as_float(g[i]) : 0.f;
              if (ext) dp += ext[kk];
              delta += pr * dp;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(sdp_empty);
        continue;
      }
      // this warp's half of the S / dP rows into registers in one go, then TMEM is free for the next tile's MMAs
      constexpr int CW = KVT / 2;
      const int cb = grp * CW;
      uint32_t s[CW], g[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 32) {
        tmem_ld_x32(tS + lane_off + cb + c, s + c);
        if (p.has_dO) tmem_ld_x32(tdP + lane_off + cb + c, g + c);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(sdp_empty);
      const bool full = ok && (kbase + KVT <= p.nk) && (ext == nullptr);
      mbar_wait(ds_empty, (j & 1) ^ 1);
#pragma unroll
      for (int c = 0; c < CW; c += 32) {
        const int c0 = cb + c;
        uint32_t pk[16];
        if (full) {       // interior tile, no loss gradient: pure arithmetic
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[c + 2 * i]), p.scale_log2, -L2));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[c + 2 * i + 1]), p.scale_log2, -L2));
            const float d0 = (p.has_dO ? __uint_as_float(g[c + 2 * i]) : 0.f) - delta;
            const float d1 = (p.has_dO ? __uint_as_float(g[c + 2 * i + 1]) : 0.f) - delta;
            pk[i] = pack_h2(p0 * d0 * p.scale, p1 * d1 * p.scale);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float o2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int kk = kbase + c0 + 2 * i + e;
              float v = 0.f;
              if (ok && kk < p.nk) {
                const float pr = ex2_approx(fmaf(__uint_as_float(s[c + 2 * i + e]), p.scale_log2, -L2));
                float dp = p.has_dO ? __uint_as_float(g[c + 2 * i + e]) : 0.f;
                if (ext) dp += ext[kk];
                v = pr * (dp - delta) * p.scale;
              }
              o2[e] = v;
            }
            pk[i] = pack_h2(o2[0], o2[1]);
          }
        }
        uint8_t* atom = sdS + (c0 >> 6) * 16384;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0 + q)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    const int b = bh / p.heads, h = bh % p.heads;
    __half* orow = p.dq + ((long long)b * p.nq + qrow) * p.ld_dq + h * p.d;
#pragma unroll 1
    for (int c0 = grp * 16; c0 < D16; c0 += 32) {   // the two warps of a quadrant interleave 16-column chunks
      uint32_t v[16];
      tmem_ld_x16(tdQ + lane_off + c0, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ dK / dV
template <int DPB, int D16, int QT, int STAGES>
struct AttnDkvCfg {
  static constexpr int K_BYTES = DPB * 16384;           // resident K and V tiles (128 rows)
  static constexpr int Q_ATOM = QT * 128;
  static constexpr int Q_BYTES = DPB * Q_ATOM;          // Q_i and dO_i tiles (QT rows)
  static constexpr int T_ATOM = D16 * 128;
  static constexpr int T_BYTES = (QT / 64) * T_ATOM;    // Q^T_i and dO^T_i tiles
  static constexpr int STAGE_BYTES = 2 * Q_BYTES + 2 * T_BYTES;
  static constexpr int PT_BYTES = (QT / 64) * 16384;
  static constexpr int SMEM_BYTES = 2 * K_BYTES + STAGES * STAGE_BYTES + 2 * PT_BYTES + 4 * QT * 4 + 1024 + 256;
};

template <int DPB, int D16, int QT, int STAGES>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmQt, const __grid_constant__ CUtensorMap tmdOt,
                    const __grid_constant__ AttnBwdParams p) {
  using Cfg = AttnDkvCfg<DPB, D16, QT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + Cfg::K_BYTES;
  uint8_t* sStage = sV + Cfg::K_BYTES;
  uint8_t* sPT = sStage + STAGES * Cfg::STAGE_BYTES;
  uint8_t* sdST = sPT + Cfg::PT_BYTES;
  float* sL = reinterpret_cast<float*>(sdST + Cfg::PT_BYTES);  // [2][QT]
  float* sD = sL + 2 * QT;                                     // [2][QT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * QT);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = q_full + STAGES;
  uint64_t* sdp_full = q_empty + STAGES;
  uint64_t* sdp_empty = sdp_full + 1;
  uint64_t* pt_full = sdp_empty + 1;
  uint64_t* pt_empty = pt_full + 1;
  uint64_t* out_full = pt_empty + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(out_full + 1);

  const int warp = threadIdx.x >> 5;
  const int kt = blockIdx.x;
  const int bh = blockIdx.y;
  const int nqt = (p.nq + QT - 1) / QT;

  if (warp == 1) {
    if (elect_one()) {
      mbar_init(kv_full, 1);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&q_full[s], 1);
        mbar_init(&q_empty[s], 1);
      }
      mbar_init(sdp_full, 1);
      mbar_init(sdp_empty, 8);
      mbar_init(pt_full, 8);
      mbar_init(pt_empty, 1);
      mbar_init(out_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr int D64 = (D16 + 63) / 64 * 64;
  const uint32_t tST = tmem_base, tdPT = tmem_base + QT, tdV = tmem_base + 2 * QT, tdK = tmem_base + 2 * QT + D64;
  static_assert(2 * QT + D64 + D16 <= 512, "TMEM budget");

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::K_BYTES);
      for (int a = 0; a < DPB; ++a) {
        tma_load_3d(sK + a * 16384, &tmK, kv_full, a * 64, kt * 128, bh);
        tma_load_3d(sV + a * 16384, &tmV, kv_full, a * 64, kt * 128, bh);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nqt; ++i) {
        mbar_wait(&q_empty[stage], phase ^ 1);
        uint8_t* sQ = sStage + stage * Cfg::STAGE_BYTES;
        uint8_t* sdO = sQ + Cfg::Q_BYTES;
        uint8_t* sQt = sdO + Cfg::Q_BYTES;
        uint8_t* sdOt = sQt + Cfg::T_BYTES;
        mbar_arrive_expect_tx(&q_full[stage], Cfg::STAGE_BYTES);
        for (int a = 0; a < DPB; ++a) {
          tma_load_3d(sQ + a * Cfg::Q_ATOM, &tmQ, &q_full[stage], a * 64, i * QT, bh);
          tma_load_3d(sdO + a * Cfg::Q_ATOM, &tmdO, &q_full[stage], a * 64, i * QT, bh);
        }
        for (int a = 0; a < QT / 64; ++a) {
          tma_load_3d(sQt + a * Cfg::T_ATOM, &tmQt, &q_full[stage], i * QT + a * 64, 0, bh);
          tma_load_3d(sdOt + a * Cfg::T_ATOM, &tmdOt, &q_full[stage], i * QT + a * 64, 0, bh);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, QT);
    constexpr uint32_t idesc_o = make_idesc_f16(128, D16);
    mbar_wait(kv_full, 0);
    int stage = 0;
    uint32_t phase = 0;
    // S^T / dP^T of query tile i+1 are issued as soon as the softmax warps hold tile i in registers (sdp_empty)
    auto issue_sdp = [&](int stage) {
      const uint32_t qaddr = smem_u32(sStage + stage * Cfg::STAGE_BYTES);
      const uint32_t oaddr = qaddr + Cfg::Q_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < DPB; ++a) {
          const uint64_t kd = make_desc_k_sw128(smem_u32(sK) + a * 16384);
          const uint64_t qd = make_desc_k_sw128(qaddr + a * Cfg::Q_ATOM);
          const uint64_t vd = make_desc_k_sw128(smem_u32(sV) + a * 16384);
          const uint64_t od = make_desc_k_sw128(oaddr + a * Cfg::Q_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ss(tST, kd + k * 2, qd + k * 2, idesc_s, (a | k) ? 1u : 0u);
            umma_f16_ss(tdPT, vd + k * 2, od + k * 2, idesc_s, (a | k) ? 1u : 0u);
          }
        }
        tc_commit(sdp_full);
      }
      __syncwarp();
    };
    mbar_wait(&q_full[0], 0);
    tc_fence_after();
    issue_sdp(0);
    for (int i = 0; i < nqt; ++i) {
      const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
      const uint32_t nphase = (nstage == 0) ? (phase ^ 1) : phase;
      auto next_sdp = [&]() {
        if (i + 1 < nqt) {
          mbar_wait(&q_full[nstage], nphase);
          mbar_wait(sdp_empty, i & 1);
          tc_fence_after();
          issue_sdp(nstage);
        }
      };
      if constexpr (STAGES >= 2) next_sdp();   // with a single stage the next tile only lands after dV/dK(i) free it
      const uint32_t qtaddr = smem_u32(sStage + stage * Cfg::STAGE_BYTES) + 2 * Cfg::Q_BYTES;
      const uint32_t otaddr = qtaddr + Cfg::T_BYTES;
      mbar_wait(pt_full, i & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < QT / 64; ++a) {
          const uint64_t pd = make_desc_k_sw128(smem_u32(sPT) + a * 16384);
          const uint64_t sd = make_desc_k_sw128(smem_u32(sdST) + a * 16384);
          const uint64_t otd = make_desc_k_sw128(otaddr + a * Cfg::T_ATOM);
          const uint64_t qtd = make_desc_k_sw128(qtaddr + a * Cfg::T_ATOM);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ss(tdV, pd + k * 2, otd + k * 2, idesc_o, (i | a | k) ? 1u : 0u);
            umma_f16_ss(tdK, sd + k * 2, qtd + k * 2, idesc_o, (i | a | k) ? 1u : 0u);
          }
        }
        tc_commit(&q_empty[stage]);
        tc_commit(pt_empty);
        if (i == nqt - 1) tc_commit(out_full);
      }
      __syncwarp();
      if constexpr (STAGES < 2) next_sdp();
      stage = nstage;
      phase = nphase;
    }
  } else {
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;   // column half (query columns) of this warp
    const int r = quad * 32 + lane_id();
    const int tid = threadIdx.x - 64;  // 0..255 among the softmax warps
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int krow = kt * 128 + r;
    const bool kok = krow < p.nk;
    // row statistic / delta of the query tile: double-buffered in shared memory, the next tile's values are fetched
    // (global loads in flight) while this tile is processed
    if (tid < QT) {
      sL[tid] = tid < p.nq ? p.lse2[(long long)bh * p.nq_alloc + tid] : 0.f;
      sD[tid] = tid < p.nq ? p.delta[(long long)bh * p.nq_alloc + tid] : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int i = 0; i < nqt; ++i) {
      const float* L = sL + (i & 1) * QT;
      const float* Dd = sD + (i & 1) * QT;
      const bool pf = (i + 1 < nqt) && (tid < QT);
      float nl = 0.f, nd = 0.f;
      if (pf) {
        const int q = (i + 1) * QT + tid;
        if (q < p.nq) {
          nl = p.lse2[(long long)bh * p.nq_alloc + q];
          nd = p.delta[(long long)bh * p.nq_alloc + q];
        }
      }
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
      // this warp's half of the S^T / dP^T rows into registers, then TMEM is free for the next query tile's MMAs
      constexpr int CW = QT / 2;
      const int cb = grp * CW;
      uint32_t s[CW], g[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 32) {
        tmem_ld_x32(tST + lane_off + cb + c, s + c);
        tmem_ld_x32(tdPT + lane_off + cb + c, g + c);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(sdp_empty);
      const bool full = kok && ((i + 1) * QT <= p.nq);
      mbar_wait(pt_empty, (i & 1) ^ 1);
#pragma unroll
      for (int cc = 0; cc < CW; cc += 32) {
        const int c0 = cb + cc;
        uint32_t pp[16], ds[16];
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {       // four query columns per step: one 16-byte read of each table
          const float4 l4 = *reinterpret_cast<const float4*>(L + c0 + 4 * e4);
          const float4 d4 = *reinterpret_cast<const float4*>(Dd + c0 + 4 * e4);
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
          float pr[4], dsv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + 4 * e4 + e;
            const float pe = ex2_approx(fmaf(__uint_as_float(s[cc + 4 * e4 + e]), p.scale_log2, -lv[e]));
            const float de = pe * (__uint_as_float(g[cc + 4 * e4 + e]) - dv[e]) * p.scale;
            const bool valid = full || (kok && (i * QT + c) < p.nq);   // ragged edges: select, no branch
            pr[e] = valid ? pe : 0.f;
            dsv[e] = valid ? de : 0.f;
          }
          pp[2 * e4] = pack_h2(pr[0], pr[1]);
          pp[2 * e4 + 1] = pack_h2(pr[2], pr[3]);
          ds[2 * e4] = pack_h2(dsv[0], dsv[1]);
          ds[2 * e4 + 1] = pack_h2(dsv[2], dsv[3]);
        }
        const int ch0 = (c0 & 63) >> 3;
        uint8_t* pa = sPT + (c0 >> 6) * 16384;
        uint8_t* da = sdST + (c0 >> 6) * 16384;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(pa + sw128_offset(r, ch0 + q)) =
              make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
          *reinterpret_cast<uint4*>(da + sw128_offset(r, ch0 + q)) =
              make_uint4(ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(pt_full);
      if (pf) {
        sL[((i + 1) & 1) * QT + tid] = nl;
        sD[((i + 1) & 1) * QT + tid] = nd;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // next tile's tables visible; this tile's reads are done
    }
    mbar_wait(out_full, 0);
    tc_fence_after();
    const int b = bh / p.heads, h = bh % p.heads;
    const bool st_ok = krow < p.nk_store;
#pragma unroll 1
    for (int which = grp; which < 2; which += 2) {   // warp group 0 drains dV, group 1 drains dK
      __half* base = which ? p.dk : p.dv;
      const int ld = which ? p.ld_dk : p.ld_dv;
      const uint32_t t = which ? tdK : tdV;
      if (!base) continue;
      __half* orow = base + ((long long)b * p.nk_store + krow) * ld + h * p.d;
#pragma unroll 1
      for (int c0 = 0; c0 < D16; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(t + lane_off + c0, v);
        tmem_ld_wait();
        if (st_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g)
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// delta[bh, q] = sum_j dO[b*nq+q, h*d+j] * O[b*nq+q, h*d+j]; one warp per (row, head)
__global__ void attn_delta_kernel(const __half* __restrict__ dO, int ld_do, const __half* __restrict__ O, int ld_o,
                                  float* __restrict__ delta, int B, int heads, int nq, int nq_alloc, int d) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)B * nq * heads;
  for (long long w = blockIdx.x * (long long)warps + (threadIdx.x >> 5); w < total; w += (long long)gridDim.x * warps) {
    const int h = (int)(w % heads);
    const long long row = w / heads;
    const int b = (int)(row / nq), q = (int)(row % nq);
    float acc = 0.f;
    for (int j = lane; j < d; j += 32)
      acc += __half2float(dO[row * ld_do + h * d + j]) * __half2float(O[row * ld_o + h * d + j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) delta[((long long)b * heads + h) * nq_alloc + q] = acc;
  }
}

}  // namespace b200
