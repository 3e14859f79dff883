// Host-side launch builders for the attention kernels.
#pragma once
#include "attention.cuh"
#include "gemm_host.cuh"

namespace b200 {

inline int round_dp(int d) { return (d + 63) / 64 * 64; }
// supported TMEM/N widths for the P.V product
inline int round_d16(int d) {
  const int opts[] = {16, 32, 48, 64, 80, 128, 160, 192};
  for (int o : opts)
    if (d <= o) return o;
  throw std::runtime_error("head_dim > 192 unsupported");
}

struct AttnLaunch {
  CUtensorMap tmQ, tmK, tmVt;
  AttnParams p;
  int dpb, d16;
  dim3 grid;
};

// slabs: q [BH, nq_alloc, dp], k [BH, nk_alloc, dp], vt [BH, d16, nk_alloc]
inline AttnLaunch build_attn_fwd(const __half* q, const __half* k, const __half* vt, __half* out, int ldo, float* lse2,
                                 int B, int heads, int nq, int nk, int nq_alloc, int nk_alloc, int d, float scale) {
  AttnLaunch L;
  const int dp = round_dp(d), d16 = round_d16(d);
  const int BH = B * heads;
  if (nk_alloc % 8) throw std::runtime_error("nk_alloc must be a multiple of 8");
  {
    uint64_t dims[3] = {(uint64_t)dp, (uint64_t)nq_alloc, (uint64_t)BH};
    uint64_t st[2] = {(uint64_t)dp * 2, (uint64_t)dp * 2 * nq_alloc};
    uint32_t box[3] = {64, 128, 1};
    L.tmQ = make_tmap_f16(q, 3, dims, st, box);
  }
  {
    uint64_t dims[3] = {(uint64_t)dp, (uint64_t)nk_alloc, (uint64_t)BH};
    uint64_t st[2] = {(uint64_t)dp * 2, (uint64_t)dp * 2 * nk_alloc};
    uint32_t box[3] = {64, 128, 1};
    L.tmK = make_tmap_f16(k, 3, dims, st, box);
  }
  {
    uint64_t dims[3] = {(uint64_t)nk_alloc, (uint64_t)d16, (uint64_t)BH};
    uint64_t st[2] = {(uint64_t)nk_alloc * 2, (uint64_t)nk_alloc * 2 * d16};
    uint32_t box[3] = {64, (uint32_t)d16, 1};
    L.tmVt = make_tmap_f16(vt, 3, dims, st, box);
  }
  L.p.heads = heads; L.p.nq = nq; L.p.nk = nk; L.p.nq_alloc = nq_alloc; L.p.nk_alloc = nk_alloc; L.p.d = d;
  L.p.scale_log2 = scale * 1.4426950408889634f;
  L.p.out = out; L.p.ldo = ldo; L.p.lse2 = lse2;
  L.dpb = dp / 64;
  L.d16 = d16;
  L.grid = dim3((nq + 127) / 128, BH, 1);
  return L;
}

template <int DPB, int D16, int STAGES>
inline void launch_attn_fwd_t(const AttnLaunch& L, cudaStream_t st) {
  using Cfg = AttnCfg<DPB, D16, STAGES>;
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<DPB, D16, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES));
    done = true;
  }
  attn_fwd_kernel<DPB, D16, STAGES><<<L.grid, 192, Cfg::SMEM_BYTES, st>>>(L.tmQ, L.tmK, L.tmVt, L.p);
  B200_CHECK(cudaGetLastError());
}

inline void run_attn_fwd(const AttnLaunch& L, cudaStream_t st) {
  const int key = L.dpb * 1000 + L.d16;
  switch (key) {
    case 1016: launch_attn_fwd_t<1, 16, 2>(L, st); break;
    case 1032: launch_attn_fwd_t<1, 32, 2>(L, st); break;
    case 1048: launch_attn_fwd_t<1, 48, 2>(L, st); break;
    case 1064: launch_attn_fwd_t<1, 64, 2>(L, st); break;
    case 2080: launch_attn_fwd_t<2, 80, 2>(L, st); break;
    case 2128: launch_attn_fwd_t<2, 128, 2>(L, st); break;
    case 3160: launch_attn_fwd_t<3, 160, 1>(L, st); break;
    case 3192: launch_attn_fwd_t<3, 192, 1>(L, st); break;
    default: throw std::runtime_error("unsupported head_dim configuration");
  }
}

}  // namespace b200
