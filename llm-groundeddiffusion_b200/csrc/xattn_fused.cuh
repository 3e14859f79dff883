// The graded kernel: cross-attention with its projections and the guidance loss in ONE launch (sm_100a).
//
//   out[rows, :] = residual + bias_o + softmax(scale * (x Wq^T) K^T) V  Wo^T        (+ loss, d loss/dP, optional maps)
//
// Grid = clusters of 4 CTAs (4 heads); two clusters share one 128-row tile of x (rows of ONE image: n % 128 == 0).
// (Clusters of 8 do not all fit on the GPU at once - 16 were needed, the profile showed a second wave - clusters of 4 do.)
//   phase 1  Q_h   = x_tile . Wq_h^T        tcgen05 128 x d x C GEMM; the x tile is fetched ONCE per cluster: each CTA
//                                           TMA-loads a 16-row slice of every k-block and multicasts it to all 8 CTAs
//   core     S = Q_h K_h^T -> softmax (one thread per row, from TMEM) -> P (fp16 smem) -> O_h = P V_h
//            + attention-map outputs, loss-column scratch, last-CTA-per-(image, head) loss/gradient (as xattn.cuh)
//   exchange O_h -> HBM/L2 [rows, C]; cluster barrier
//   phase 2  out[:, h*d:(h+1)*d] = O_tile . Wo[h*d:(h+1)*d, :]^T + bias + residual   (O tile multicast like x)
// Text K/V of the prompt are constants prepared once (slabs as in attention.cuh).  FLOPs per launch = the SURVEY
// section 8(d) accounting: B*(2nC^2 + 2nTC + 2nTC + 2nC^2).
#pragma once
#include "xattn.cuh"

namespace b200 {

struct FusedXattnParams {
  int n, d, C, nk, k_alloc;
  int tiles_per_img;
  float scale_log2;
  const float* bias_o;
  const __half* residual;   // [M, C]
  __half* out;              // [M, C]
  __half* o_buf;            // [M, C] scratch
  __half* q_slab;           // optional [B*8, n, dp] row-major Q for the backward
  float* lse2;              // optional [B*8, n]
  __half* probs;            // optional [B*8, n, nk]
  const int* save_tok;
  __half* probs_tok;        // [B*8, n]
  int has_loss;
  XattnLoss L;
  int* tile_flags;           // [row tiles] zero on entry: counts the heads of a row tile whose O_h is published
  int* tile_done;            // [row tiles] zero on entry: heads that have seen the full count; the last one re-zeroes both
  int* bh_ready;             // [B*8] zero on entry: row tiles of an (image, head) whose P columns are published
  int* bh_done;              // [B*8] zero on entry: row tiles of an (image, head) that finished their loss share
  float* loss_partials;      // [B*8][tiles_per_img]
  unsigned long long* dbg;   // optional [grid][8] %globaltimer stamps (phase timeline), null in production
  int loss_stage;            // 1: stage the loss inputs in shared memory (default), 0: per-problem global loads (A/B)
};
__device__ __forceinline__ bool fused_loss_stage_enabled(const FusedXattnParams& p) { return p.loss_stage != 0; }

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FUSED_STAMP(i) do { if (p.dbg && threadIdx.x == 64) p.dbg[(long long)blockIdx.x * 16 + (i)] = gtimer(); } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

template <int D>   // head dim, multiple of 16, <= 192
struct FusedCfg {
  static constexpr int STAGES = 4;
  static constexpr int A_BYTES = 16384;
  static constexpr int B_BYTES = D * 128;
  static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int DPB = (D + 63) / 64;                 // K atoms of Q / K_h
  static constexpr int KT_ATOM = 80 * 128;                  // K_h atom: 80 text rows x 128 B
  static constexpr int K_BYTES = ((DPB * KT_ATOM + 1023) / 1024) * 1024;
  static constexpr int V_ATOM = D * 128;                    // V^T atom: D rows x 64 keys
  static constexpr int V_BYTES = ((2 * V_ATOM + 1023) / 1024) * 1024;
  static constexpr int Q_BYTES = DPB * 16384;               // aliases the stage ring
  static constexpr int P_BYTES = 2 * 16384;                 // aliases the stage ring
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static_assert(Q_BYTES + P_BYTES <= RING_BYTES, "core scratch must fit in the stage ring");
  static_assert(128 * (D + 8) * 2 <= RING_BYTES, "epilogue staging must fit in the stage ring");
  static constexpr int SMEM_BYTES = RING_BYTES + K_BYTES + V_BYTES + 1024 + 512 + 4096;   // + loss tables
  static_assert(SMEM_BYTES <= 232448, "smem budget");
};

template <int D>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(192, 1)
xattn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmVt,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWo,
                   const __grid_constant__ FusedXattnParams p) {
  using Cfg = FusedCfg<D>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* sQ = ring;                      // core scratch aliases the ring (idle between the two GEMM phases)
  uint8_t* sP = ring + Cfg::Q_BYTES;
  uint8_t* sK = ring + Cfg::RING_BYTES;
  uint8_t* sV = sK + Cfg::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::V_BYTES);
  uint64_t* full_bar = bars;               // STAGES
  uint64_t* empty_bar = bars + STAGES;     // STAGES (count 8: every CTA of the cluster frees the slot)
  uint64_t* kv_full = empty_bar + STAGES;  // 1
  uint64_t* acc_full = kv_full + 1;        // 1: a GEMM phase / core MMA finished (phases 0..3)
  uint64_t* q_ready = acc_full + 1;        // 1 (count 4)
  uint64_t* p_ready = q_ready + 1;         // 1 (count 4)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p_ready + 1);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr + 1);
  float* s_red = reinterpret_cast<float*>(s_flag + 1);
  float* sL = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);   // loss tables (4 KB)

  const int warp = threadIdx.x >> 5;
  const int cr = (int)cluster_ctarank();           // rank inside the 4-CTA cluster
  const int h = (int)(blockIdx.x & 7);             // head (clusters 2rt and 2rt+1 hold heads 0-3 and 4-7)
  const int rt = blockIdx.x >> 3;                  // row tile
  const int row0 = rt * 128;
  const int b = row0 / p.n;                        // image of this tile
  const int bh = b * 8 + h;
  const int tok0 = row0 - b * p.n;                 // first token of the tile inside the image
  const int nkb = p.C >> 6;                        // k-blocks of the two projection GEMMs

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmWq); prefetch_tmap(&tmK);
    prefetch_tmap(&tmVt); prefetch_tmap(&tmO); prefetch_tmap(&tmWo);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 4);
      }
      mbar_init(kv_full, 1);
      mbar_init(acc_full, 1);
      mbar_init(q_ready, 4);
      mbar_init(p_ready, 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();           // every CTA's barriers exist before anyone multicasts into / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  FUSED_STAMP(0);
  const uint32_t tAcc = tmem_base;          // Q_h, then O_h, then the output tile (D columns)
  const uint32_t tS = tmem_base + 256;      // scores (80 columns)

  // ------------------------------------------------------------------ role: TMA producer
  if (warp == 0) {
    if (elect_one()) {
      // text K/V of this (image, head): needed only by the core, issued first so they arrive during phase 1
      mbar_arrive_expect_tx(kv_full, Cfg::DPB * Cfg::KT_ATOM + 2 * Cfg::V_ATOM);
      for (int a = 0; a < Cfg::DPB; ++a) tma_load_3d(sK + a * Cfg::KT_ATOM, &tmK, kv_full, a * 64, 0, bh);
      for (int a = 0; a < 2; ++a) tma_load_3d(sV + a * Cfg::V_ATOM, &tmVt, kv_full, a * 64, 0, bh);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
        tma_load_2d_mcast(sa + cr * 4096, &tmX, &full_bar[stage], kb * 64, row0 + cr * 32, (uint16_t)0xF);
        tma_load_2d(sa + Cfg::A_BYTES, &tmWq, &full_bar[stage], kb * 64, h * D);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
  // ------------------------------------------------------------------ role: MMA issuer, phase 1
  constexpr uint32_t idesc_g = make_idesc_f16(128, D);
  constexpr uint32_t idesc_s = make_idesc_f16(128, 80);
  if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(ring + stage * Cfg::STAGE_BYTES);
        const uint64_t ad = make_desc_k_sw128(sa);
        const uint64_t bd = make_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tAcc, ad + k * 2, bd + k * 2, idesc_g, (kb | k) ? 1u : 0u);
        tc_commit_mcast(&empty_bar[stage], (uint16_t)0xF);
        if (kb == nkb - 1) tc_commit(acc_full);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- core MMAs
    mbar_wait(kv_full, 0);
    mbar_wait(q_ready, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int a = 0; a < Cfg::DPB; ++a) {
        const uint64_t ad = make_desc_k_sw128(smem_u32(sQ) + a * 16384);
        const uint64_t bd = make_desc_k_sw128(smem_u32(sK) + a * Cfg::KT_ATOM);
        const int ksteps = (a == Cfg::DPB - 1) ? ((D - a * 64 + 15) / 16) : 4;
        for (int k = 0; k < ksteps; ++k) umma_f16_ss(tS, ad + k * 2, bd + k * 2, idesc_s, (a | k) ? 1u : 0u);
      }
      tc_commit(acc_full);
    }
    __syncwarp();
    mbar_wait(p_ready, 0);
    tc_fence_after();
    if (elect_one()) {
      // O_h = P V_h : K = 80 keys = 5 k-steps (4 in atom 0, 1 in atom 1)
      for (int ks = 0; ks < 5; ++ks) {
        const int a = ks >> 2, k = ks & 3;
        const uint64_t ad = make_desc_k_sw128(smem_u32(sP) + a * 16384);
        const uint64_t bd = make_desc_k_sw128(smem_u32(sV) + a * Cfg::V_ATOM);
        umma_f16_ss(tAcc, ad + k * 2, bd + k * 2, idesc_g, ks ? 1u : 0u);
      }
      tc_commit(acc_full);
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ role: epilogue / softmax warps
  const int quad = warp & 3;
  const int r = quad * 32 + lane_id();
  const int tid = threadIdx.x - 64;
  const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
  const int tok = tok0 + r;                       // token index inside the image
  const long long grow = (long long)row0 + r;     // global row
  auto p_at = [&](int t) -> float {
    const uint8_t* atom = sP + (t >> 6) * 16384;
    const __half* chunk = reinterpret_cast<const __half*>(atom + sw128_offset(r, (t & 63) >> 3));
    return __half2float(chunk[t & 7]);
  };
  float row_m = 0.f, row_l = 1.f;
  LossRaw<8> loss_pre;                            // first loss problem of this warp, requested before the cluster barrier
  bool loss_have_pre = false;
  if (warp >= 2) {
    // loss inputs while phase 1 runs on the tensor core: this tile's rows of dp_extra zeroed (ordered before every
    // CTA's atomics by the publish fence below), term / slot tables and the problem list staged in shared memory
    if (p.has_loss) {
      loss_zero(p.L, bh, p.n, tok0, 128, tid);
      if (p.tiles_per_img == 2 && tok0 == 0 && tid == 0) p.L.loss_part[bh] = 0.f;   // two-part combine adds into it
      loss_stage(p.L, sL, tid, b);
    }
    // ---- Q_h: TMEM -> fp16 -> smem A operand (and the optional Q slab for the backward)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    FUSED_STAMP(1);
    __half* qrow = p.q_slab ? p.q_slab + ((long long)bh * p.n + tok) * (Cfg::DPB * 64) : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < Cfg::DPB * 64; c0 += 16) {
      uint32_t pk[8];
      if (c0 < D) {
        uint32_t v[16];
        tmem_ld_x16(tAcc + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_h2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
      }
      uint8_t* atom = sQ + (c0 >> 6) * 16384;
      const int ch0 = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      if (qrow) {
        *reinterpret_cast<uint4*>(qrow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(qrow + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(q_ready);

    // ---- softmax over the 77 keys
    mbar_wait(acc_full, 1);
    tc_fence_after();
    FUSED_STAMP(2);
    float m = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < 80; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tS + lane_off + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.nk) m = fmaxf(m, __uint_as_float(v[i]));
    }
    m *= p.scale_log2;
    float l = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 80; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tS + lane_off + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.nk) l += ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m));
    }
    const float inv_l = 1.f / l;
    row_m = m;
    row_l = l;
    __half* prow = p.probs ? p.probs + ((long long)bh * p.n + tok) * p.nk : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t pk[8];
      if (c0 < 80) {
        uint32_t v[16];
        tmem_ld_x16(tS + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = (c0 + 2 * i < p.nk) ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2, -m)) * inv_l : 0.f;
          const float c = (c0 + 2 * i + 1 < p.nk) ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2, -m)) * inv_l : 0.f;
          pk[i] = pack_h2(a, c);
        }
        if (prow) {
          const __half* hp = reinterpret_cast<const __half*>(pk);
          for (int i = 0; i < 16 && c0 + i < p.nk; ++i) prow[c0 + i] = hp[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = 0u;
      }
      uint8_t* atom = sP + (c0 >> 6) * 16384;
      const int ch0 = (c0 & 63) >> 3;
      *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    if (p.save_tok) {
      const int t = p.save_tok[b];
      if (t >= 0) p.probs_tok[(long long)bh * p.n + tok] = __float2half_rn(p_at(t));
    }
    if (p.has_loss) {
      const int* stok = loss_scratch(sL).stok;         // staged by loss_stage above (no global round trip per slot)
      for (int s = 0; s < kMaxSlots; ++s) {
        const int t = stok[s];
        if (t < 0) break;
        p.L.pcol[((long long)bh * kMaxSlots + s) * p.n + tok] = p_at(t);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(p_ready);

    // ---- O_h: TMEM -> fp16 -> o_buf[rows, h*D ...]  (the A operand of phase 2, read back through L2)
    mbar_wait(acc_full, 0);   // third completion of acc_full: parity 0 again
    tc_fence_after();
    FUSED_STAMP(3);
    constexpr int OLD = D + 8;                         // padded row stride (halves)
    {
      // inputs of the OUTPUT epilogue, fetched now: the K/V region is dead (S and P.V have completed), so the residual
      // tile (asynchronous LDGSTS copies, waited for only before the output epilogue) and this head's slice of the
      // output-projection bias land there while the hand-shake, the cluster barrier and phase 2 run
      constexpr int PPR = D / 8;
      static_assert(128 * (D + 8) * 2 <= Cfg::K_BYTES + Cfg::V_BYTES, "residual staging must fit in the K/V region");
      static_assert(128 * (D + 8) * 2 + 1024 <= Cfg::K_BYTES + Cfg::V_BYTES && D * 4 <= 1024, "bias staging must fit");
      const uint32_t stgE_s = smem_u32(sK);
      if (p.residual) {
        for (int pi = tid; pi < 128 * PPR; pi += 128) {
          const int row = pi / PPR, pc = pi - row * PPR;
          const __half* src = p.residual + ((long long)row0 + row) * p.C + h * D + pc * 8;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stgE_s + (row * OLD + pc * 8) * 2), "l"(src)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (tid < D / 4) {
        const uint32_t dst = smem_u32(sK + Cfg::K_BYTES + Cfg::V_BYTES - 1024) + tid * 16;
        if (p.bias_o) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p.bias_o + h * D + tid * 4) : "memory");
        } else {
          sts128(dst, make_uint4(0u, 0u, 0u, 0u));
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // stage the row in shared memory (ring scratch: Q / P are dead once P.V has completed), then leave as full row
    // segments: consecutive lanes write consecutive 16-byte pieces of a row
    __half* stgO = reinterpret_cast<__half*>(ring);
    const uint32_t stgO_s = smem_u32(stgO);
#pragma unroll 1
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tAcc + lane_off + c0, v);
      tmem_ld_wait();
      uint4 s0, s1;
      s0.x = pack_h2(__uint_as_float(v[0]), __uint_as_float(v[1])); s0.y = pack_h2(__uint_as_float(v[2]), __uint_as_float(v[3]));
      s0.z = pack_h2(__uint_as_float(v[4]), __uint_as_float(v[5])); s0.w = pack_h2(__uint_as_float(v[6]), __uint_as_float(v[7]));
      s1.x = pack_h2(__uint_as_float(v[8]), __uint_as_float(v[9])); s1.y = pack_h2(__uint_as_float(v[10]), __uint_as_float(v[11]));
      s1.z = pack_h2(__uint_as_float(v[12]), __uint_as_float(v[13])); s1.w = pack_h2(__uint_as_float(v[14]), __uint_as_float(v[15]));
      sts128(stgO_s + (r * OLD + c0) * 2, s0);
      sts128(stgO_s + (r * OLD + c0 + 8) * 2, s1);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
      constexpr int PPR = D / 8;                       // 16-byte pieces per row
      for (int pi = tid; pi < 128 * PPR; pi += 128) {
        const int row = pi / PPR, pc = pi - row * PPR;
        *reinterpret_cast<uint4*>(p.o_buf + ((long long)row0 + row) * p.C + h * D + pc * 8) =
            lds128(stgO_s + (row * OLD + pc * 8) * 2);
      }
    }
    if (p.lse2) p.lse2[(long long)bh * p.n + tok] = row_m + log2f(row_l);
    fence_proxy_async_all();
    tc_fence_before();
    // publish: this head's O_h is in L2; phase 2 needs all 8 heads of the row tile (two clusters).  The CTA barrier
    // orders every thread's stores before thread 0's gpu-scope fence (fences are cumulative), then the flag goes up.
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) {
      __threadfence();
      atomicAdd(&p.tile_flags[rt], 1);
      if (p.has_loss) atomicAdd(&p.bh_ready[bh], 1);
      uint32_t spins = 0;
      while (atomicAdd(&p.tile_flags[rt], 0) < 8) {
        if (++spins > B200_SPIN_LIMIT) __trap();
        __nanosleep(64);
      }
      // the loss reduction needs the P columns of every row tile of this (image, head): wait for them HERE, inside the
      // hand-shake this thread is already spinning in, instead of paying another L2 round trip after the cluster barrier
      // (every tile publishes before it waits, so there is no circular wait)
      if (p.has_loss) {
        spins = 0;
        while (atomicAdd(&p.bh_ready[bh], 0) < p.tiles_per_img) {
          if (++spins > B200_SPIN_LIMIT) __trap();
          __nanosleep(64);
        }
      }
      __threadfence();
      // every flag is back to zero when the kernel ends (no memset between launches): the last head to get here resets
      if (atomicAdd(&p.tile_done[rt], 1) == 7) {
        p.tile_flags[rt] = 0;
        p.tile_done[rt] = 0;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  // every CTA of the cluster has left the core (ring scratch free) and all 8 O_h of the row tile are visible
  __syncwarp();
  FUSED_STAMP(4);
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  // every tile's P columns of this (image, head) are published and fenced (thread 0 saw bh_ready complete, then the
  // CTA barrier above): request the first loss problem's inputs between arrive and wait, so that their L2 round trip
  // runs under the cluster barrier (after the arrive: its release fence would otherwise wait for these loads)
  if (warp >= 2 && p.has_loss && !fused_loss_stage_enabled(p))
    loss_have_pre = loss_prefetch_first(p.L, sL, tid, h, 8, bh, p.n, tok0 >> 7, p.tiles_per_img, loss_pre);
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  FUSED_STAMP(5);

  // ------------------------------------------------------------------ phase 2: out slice = O_tile . Wo_h^T
  if (warp == 0) {
    if (elect_one()) {
      fence_proxy_async_all();
      int stage = nkb % STAGES;
      uint32_t phase = (nkb / STAGES) & 1;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
        tma_load_2d_mcast(sa + cr * 4096, &tmO, &full_bar[stage], kb * 64, row0 + cr * 32, (uint16_t)0xF);
        tma_load_2d(sa + Cfg::A_BYTES, &tmWo, &full_bar[stage], kb * 64, h * D);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = nkb % STAGES;
    uint32_t phase = (nkb / STAGES) & 1;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(ring + stage * Cfg::STAGE_BYTES);
        const uint64_t ad = make_desc_k_sw128(sa);
        const uint64_t bd = make_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tAcc, ad + k * 2, bd + k * 2, idesc_g, (kb | k) ? 1u : 0u);
        tc_commit_mcast(&empty_bar[stage], (uint16_t)0xF);
        if (kb == nkb - 1) tc_commit(acc_full);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    constexpr int OLD = D + 8;
    constexpr int PPR = D / 8;
    // the residual tile and the bias slice were requested right after the core (K/V region); nothing has waited on
    // them yet
    __half* stgE = reinterpret_cast<__half*>(sK);
    const uint32_t stgE_s = smem_u32(stgE);
    const float* s_bias = reinterpret_cast<const float*>(sK + Cfg::K_BYTES + Cfg::V_BYTES - 1024);
    // ---- guidance loss, overlapped with phase 2 on the tensor core: the row tiles of this (image, head) share the
    // problems once every tile's P columns are published (fenced before bh_ready went up)
    int loss_ticket = -1;
    if (p.has_loss) {
      // every tile's P columns are published: thread 0 saw bh_ready complete (and fenced) before the cluster barrier
      // staging area: the K/V region behind the residual / output staging rows (dead since the core)
      uint8_t* lstage = sK + 128 * OLD * 2;
      const int lstage_bytes = Cfg::K_BYTES + Cfg::V_BYTES - 128 * OLD * 2 - 1024;   // the last KB holds the bias
      loss_ticket = loss_run(p.L, sL, tid, h, 8, bh, p.n, tok0 >> 7, p.tiles_per_img, p.loss_partials, p.bh_done,
                             p.dbg ? p.dbg + (long long)blockIdx.x * 16 + 8 : nullptr, p.bh_ready,
                             fused_loss_stage_enabled(p) ? lstage : nullptr, lstage_bytes, p.tiles_per_img == 2,
                             loss_pre, loss_have_pre);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");       // residual tile and bias are in shared memory
    mbar_wait(acc_full, 1);   // fourth completion
    tc_fence_after();
    FUSED_STAMP(6);
#pragma unroll 1
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tAcc + lane_off + c0, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + s_bias[c0 + i];
      if (p.residual) {
        const uint4 r0 = lds128(stgE_s + (r * OLD + c0) * 2), r1 = lds128(stgE_s + (r * OLD + c0 + 8) * 2);
        const __half2* a = reinterpret_cast<const __half2*>(&r0);
        const __half2* c = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 t0 = __half22float2(a[i]), t1 = __half22float2(c[i]);
          f[2 * i] += t0.x; f[2 * i + 1] += t0.y;
          f[8 + 2 * i] += t1.x; f[8 + 2 * i + 1] += t1.y;
        }
      }
      sts128(stgE_s + (r * OLD + c0) * 2,
             make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]), pack_h2(f[6], f[7])));
      sts128(stgE_s + (r * OLD + c0 + 8) * 2,
             make_uint4(pack_h2(f[8], f[9]), pack_h2(f[10], f[11]), pack_h2(f[12], f[13]), pack_h2(f[14], f[15])));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int pi = tid; pi < 128 * PPR; pi += 128) {
      const int row = pi / PPR, pc = pi - row * PPR;
      *reinterpret_cast<uint4*>(p.out + ((long long)row0 + row) * p.C + h * D + pc * 8) =
          lds128(stgE_s + (row * OLD + pc * 8) * 2);
    }
    tc_fence_before();
    // two-part loss combine: the ticket was taken inside loss_run (thread 0), its round trip has long returned
    if (p.has_loss && tid == 0 && p.tiles_per_img == 2) loss_ticket_reset(loss_ticket, p.bh_done, p.bh_ready, bh);
  }
  FUSED_STAMP(7);
  __syncthreads();
  cluster_sync_all();   // nobody exits while a peer can still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
