// Host-side launch builders for the attention kernels.
#pragma once
#include "attention.cuh"
#include "attention2.cuh"
#include "attention_bwd.cuh"
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

inline bool& attn_use_v2() {
  static bool v = true;   // second-generation kernel (attention2.cuh) for head_dim <= 64
  return v;
}

inline void run_attn_fwd(const AttnLaunch& L, cudaStream_t st) {
  const int key = L.dpb * 1000 + L.d16;
  if (L.dpb == 1 && attn_use_v2()) {
    const int BH = (int)L.grid.y;
    switch (L.d16) {
      case 16: launch_attn_fwd2_t<16>(L.tmQ, L.tmK, L.tmVt, L.p, L.p.nq, BH, st); break;
      case 32: launch_attn_fwd2_t<32>(L.tmQ, L.tmK, L.tmVt, L.p, L.p.nq, BH, st); break;
      case 48: launch_attn_fwd2_t<48>(L.tmQ, L.tmK, L.tmVt, L.p, L.p.nq, BH, st); break;
      default: launch_attn_fwd2_t<64>(L.tmQ, L.tmK, L.tmVt, L.p, L.p.nq, BH, st); break;
    }
    B200_CHECK(cudaGetLastError());
    return;
  }
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


// ------------------------------------------------------------------------------------------------ backward
inline CUtensorMap slab_rm_map(const __half* p, int BH, int n_alloc, int dp, int box_rows) {
  uint64_t dims[3] = {(uint64_t)dp, (uint64_t)n_alloc, (uint64_t)BH};
  uint64_t st[2] = {(uint64_t)dp * 2, (uint64_t)dp * 2 * n_alloc};
  uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tmap_f16(p, 3, dims, st, box);
}
inline CUtensorMap slab_tr_map(const __half* p, int BH, int d16, int n_alloc) {
  if (n_alloc % 8) throw std::runtime_error("transposed slab alloc must be a multiple of 8");
  uint64_t dims[3] = {(uint64_t)n_alloc, (uint64_t)d16, (uint64_t)BH};
  uint64_t st[2] = {(uint64_t)n_alloc * 2, (uint64_t)n_alloc * 2 * d16};
  uint32_t box[3] = {64, (uint32_t)d16, 1};
  return make_tmap_f16(p, 3, dims, st, box);
}

inline int dq_kvt(int dpb) { return dpb == 1 ? 128 : 64; }
inline int dkv_qt(int dpb, int d16) { return (dpb == 1 && d16 < 64) ? 128 : 64; }

struct AttnDqLaunch {
  CUtensorMap tmQ, tmdO, tmK, tmV, tmKt;
  AttnBwdParams p;
  int dpb, d16;
  dim3 grid;
};
struct AttnDkvLaunch {
  CUtensorMap tmK, tmV, tmQ, tmdO, tmQt, tmdOt;
  AttnBwdParams p;
  int dpb, d16;
  dim3 grid;
};

// q, dO: [BH, nq_alloc, dp]; k, v: [BH, nk_alloc, dp]; kt: [BH, d16, nk_alloc]
inline AttnDqLaunch build_attn_dq(const __half* q, const __half* dO, const __half* k, const __half* v,
                                  const __half* kt, const float* lse2, const float* delta, const float* extra,
                                  int ext_ld, __half* dq, int ld_dq, int B, int heads, int nq, int nk, int nq_alloc,
                                  int nk_alloc, int d, float scale) {
  AttnDqLaunch L;
  memset(&L, 0, sizeof(L));
  const int dp = round_dp(d), d16 = round_d16(d), BH = B * heads;
  L.dpb = dp / 64; L.d16 = d16;
  const int kvt = dq_kvt(L.dpb);
  L.tmQ = slab_rm_map(q, BH, nq_alloc, dp, 128);
  L.tmdO = slab_rm_map(dO ? dO : q, BH, nq_alloc, dp, 128);
  L.tmK = slab_rm_map(k, BH, nk_alloc, dp, kvt);
  L.tmV = slab_rm_map(v ? v : k, BH, nk_alloc, dp, kvt);
  L.tmKt = slab_tr_map(kt, BH, d16, nk_alloc);
  AttnBwdParams& p = L.p;
  p.heads = heads; p.nq = nq; p.nk = nk; p.nq_alloc = nq_alloc; p.nk_alloc = nk_alloc; p.d = d;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta; p.extra = extra; p.ext_ld = ext_ld; p.has_dO = dO != nullptr;
  p.dq = dq; p.ld_dq = ld_dq;
  L.grid = dim3((nq + 127) / 128, BH, 1);
  return L;
}

template <int DPB, int D16, int KVT, int STAGES>
inline void launch_dq_t(const AttnDqLaunch& L, cudaStream_t st) {
  using Cfg = AttnDqCfg<DPB, D16, KVT, STAGES>;
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<DPB, D16, KVT, STAGES>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    done = true;
  }
  attn_bwd_dq_kernel<DPB, D16, KVT, STAGES><<<L.grid, 320, Cfg::SMEM_BYTES, st>>>(L.tmQ, L.tmdO, L.tmK, L.tmV, L.tmKt,
                                                                                 L.p);
  B200_CHECK(cudaGetLastError());
}
inline void run_attn_dq(const AttnDqLaunch& L, cudaStream_t st) {
  switch (L.dpb * 1000 + L.d16) {
    case 1016: launch_dq_t<1, 16, 128, 2>(L, st); break;
    case 1032: launch_dq_t<1, 32, 128, 2>(L, st); break;
    case 1048: launch_dq_t<1, 48, 128, 2>(L, st); break;
    case 1064: launch_dq_t<1, 64, 128, 2>(L, st); break;
    case 2080: launch_dq_t<2, 80, 64, 2>(L, st); break;
    case 2128: launch_dq_t<2, 128, 64, 2>(L, st); break;
    case 3160: launch_dq_t<3, 160, 64, 1>(L, st); break;
    case 3192: launch_dq_t<3, 192, 64, 1>(L, st); break;
    default: throw std::runtime_error("unsupported head_dim configuration");
  }
}

// additionally qt, dOt: [BH, d16, nq_alloc]
inline AttnDkvLaunch build_attn_dkv(const __half* k, const __half* v, const __half* q, const __half* dO,
                                    const __half* qt, const __half* dOt, const float* lse2, const float* delta,
                                    __half* dk, int ld_dk, __half* dv, int ld_dv, int nk_store, int B, int heads,
                                    int nq, int nk, int nq_alloc, int nk_alloc, int d, float scale) {
  AttnDkvLaunch L;
  memset(&L, 0, sizeof(L));
  const int dp = round_dp(d), d16 = round_d16(d), BH = B * heads;
  L.dpb = dp / 64; L.d16 = d16;
  const int qtile = dkv_qt(L.dpb, d16);
  L.tmK = slab_rm_map(k, BH, nk_alloc, dp, 128);
  L.tmV = slab_rm_map(v, BH, nk_alloc, dp, 128);
  L.tmQ = slab_rm_map(q, BH, nq_alloc, dp, qtile);
  L.tmdO = slab_rm_map(dO, BH, nq_alloc, dp, qtile);
  L.tmQt = slab_tr_map(qt, BH, d16, nq_alloc);
  L.tmdOt = slab_tr_map(dOt, BH, d16, nq_alloc);
  AttnBwdParams& p = L.p;
  p.heads = heads; p.nq = nq; p.nk = nk; p.nq_alloc = nq_alloc; p.nk_alloc = nk_alloc; p.d = d;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2; p.delta = delta; p.has_dO = 1;
  p.dk = dk; p.ld_dk = ld_dk; p.dv = dv; p.ld_dv = ld_dv; p.nk_store = nk_store;
  L.grid = dim3((nk_store + 127) / 128, BH, 1);
  return L;
}
template <int DPB, int D16, int QT, int STAGES>
inline void launch_dkv_t(const AttnDkvLaunch& L, cudaStream_t st) {
  using Cfg = AttnDkvCfg<DPB, D16, QT, STAGES>;
  static_assert(Cfg::SMEM_BYTES <= 232448, "smem budget");
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<DPB, D16, QT, STAGES>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    done = true;
  }
  attn_bwd_dkv_kernel<DPB, D16, QT, STAGES><<<L.grid, 320, Cfg::SMEM_BYTES, st>>>(L.tmK, L.tmV, L.tmQ, L.tmdO, L.tmQt,
                                                                                 L.tmdOt, L.p);
  B200_CHECK(cudaGetLastError());
}
inline void run_attn_dkv(const AttnDkvLaunch& L, cudaStream_t st) {
  switch (L.dpb * 1000 + L.d16) {
    case 1016: launch_dkv_t<1, 16, 128, 2>(L, st); break;
    case 1032: launch_dkv_t<1, 32, 128, 2>(L, st); break;
    case 1048: launch_dkv_t<1, 48, 128, 2>(L, st); break;
    case 1064: launch_dkv_t<1, 64, 64, 2>(L, st); break;
    case 2080: launch_dkv_t<2, 80, 64, 2>(L, st); break;
    case 2128: launch_dkv_t<2, 128, 64, 2>(L, st); break;
    case 3160: launch_dkv_t<3, 160, 64, 1>(L, st); break;
    case 3192: launch_dkv_t<3, 192, 64, 1>(L, st); break;
    default: throw std::runtime_error("unsupported head_dim configuration");
  }
}

inline void run_attn_delta(const __half* dO, int ld_do, const __half* O, int ld_o, float* delta, int B, int heads,
                           int nq, int nq_alloc, int d, cudaStream_t st) {
  const long long warps = (long long)B * nq * heads;
  long long blocks = (warps + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attn_delta_kernel<<<(int)blocks, 256, 0, st>>>(dO, ld_do, O, ld_o, delta, B, heads, nq, nq_alloc, d);
  B200_CHECK(cudaGetLastError());
}

}  // namespace b200
