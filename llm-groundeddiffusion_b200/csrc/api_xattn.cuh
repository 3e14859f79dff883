// C-ABI: fused cross-attention + guidance loss forward (included by api_ops.cu)
#pragma once
#include "xattn.cuh"

namespace b200 {

struct XattnLaunch {
  CUtensorMap tmQ, tmK, tmVt;
  XattnParams p;
  int dpb, d16;
  dim3 grid;
};

inline XattnLaunch build_xattn(const __half* q, const __half* k, const __half* vt, __half* out, int ldo, float* lse2,
                               __half* probs, const int* save_tok, __half* probs_tok, const XattnLoss* loss, int B,
                               int heads, int nq, int nk, int nq_alloc, int nk_alloc, int d, float scale) {
  XattnLaunch L;
  memset(&L, 0, sizeof(L));
  if (nk > 128) throw std::runtime_error("xattn: at most 128 keys");
  if (loss && nq > 4096) throw std::runtime_error("xattn loss: at most 4096 query tokens per image");
  const int dp = round_dp(d), d16 = round_d16(d), BH = B * heads;
  L.dpb = dp / 64; L.d16 = d16;
  L.tmQ = slab_rm_map(q, BH, nq_alloc, dp, 128);
  L.tmK = slab_rm_map(k, BH, nk_alloc, dp, 128);
  L.tmVt = slab_tr_map(vt, BH, d16, nk_alloc);
  XattnParams& p = L.p;
  p.heads = heads; p.nq = nq; p.nk = nk; p.nq_alloc = nq_alloc; p.nk_alloc = nk_alloc; p.d = d;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.ldo = ldo; p.lse2 = lse2; p.probs = probs; p.save_tok = save_tok; p.probs_tok = probs_tok;
  p.has_loss = loss != nullptr;
  if (loss) p.L = *loss;
  L.grid = dim3((nq + 127) / 128, BH, 1);
  return L;
}

template <int DPB, int D16>
inline void launch_xattn_t(const XattnLaunch& L, cudaStream_t st) {
  using Cfg = XattnCfg<DPB, D16>;
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(xattn_fwd_kernel<DPB, D16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES));
    done = true;
  }
  xattn_fwd_kernel<DPB, D16><<<L.grid, 192, Cfg::SMEM_BYTES, st>>>(L.tmQ, L.tmK, L.tmVt, L.p);
  B200_CHECK(cudaGetLastError());
}
inline void run_xattn(const XattnLaunch& L, cudaStream_t st) {
  switch (L.dpb * 1000 + L.d16) {
    case 1016: launch_xattn_t<1, 16>(L, st); break;
    case 1032: launch_xattn_t<1, 32>(L, st); break;
    case 1048: launch_xattn_t<1, 48>(L, st); break;
    case 1064: launch_xattn_t<1, 64>(L, st); break;
    case 2080: launch_xattn_t<2, 80>(L, st); break;
    case 2128: launch_xattn_t<2, 128>(L, st); break;
    case 3160: launch_xattn_t<3, 160>(L, st); break;
    case 3192: launch_xattn_t<3, 192>(L, st); break;
    default: throw std::runtime_error("unsupported head_dim configuration");
  }
}
}  // namespace b200

static_assert(sizeof(b200lmd_xattn_loss) == sizeof(b200::XattnLoss), "C ABI struct mismatch");
static_assert(sizeof(b200lmd_loss_term) == sizeof(b200::LossTerm), "C ABI struct mismatch");

extern "C" int b200lmd_xattn_fwd_f16(const void* q, const void* k, const void* vt, void* out, int ldo, void* lse2,
                                     void* probs, const int* save_tok, void* probs_tok,
                                     const b200lmd_xattn_loss* loss, int B, int heads, int nq, int nk, int q_alloc,
                                     int k_alloc, int head_dim, float scale, void* stream) {
  return b200::guarded([&] {
    b200::run_xattn(b200::build_xattn((const __half*)q, (const __half*)k, (const __half*)vt, (__half*)out, ldo,
                                      (float*)lse2, (__half*)probs, save_tok, (__half*)probs_tok,
                                      reinterpret_cast<const b200::XattnLoss*>(loss), B, heads, nq, nk, q_alloc,
                                      k_alloc, head_dim, scale),
                    (cudaStream_t)stream);
  });
}
extern "C" int b200lmd_max_loss_slots(void) { return b200::kMaxSlots; }
