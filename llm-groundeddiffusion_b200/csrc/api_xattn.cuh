// C-ABI: fused cross-attention + guidance loss forward (included by api_ops.cu)
#pragma once
#include "xattn.cuh"
#include "xattn_fused.cuh"

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
  if (loss && nq > 1280) throw std::runtime_error("xattn loss: at most 1280 query tokens per image");
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

namespace b200 {
// A/B switch (b200lmd_set_option "fused_loss_stage"): staging the loss inputs in shared memory with one cooperative
// pass measured no faster than per-problem loads (profiles/r2/xattn_fused_timeline_call3.txt: 9.2 vs 8.5 us for the
// problem loop) - every L2 round trip of the epilogue warps queues behind phase 2's TMA stream either way - so it is off
inline bool& fused_loss_stage() {
  static bool on = false;
  return on;
}
inline unsigned long long*& fused_dbg() {
  static unsigned long long* p = nullptr;
  return p;
}
// hand-shake counters + per (image, head, row tile) loss partials of xattn_fused_kernel.  One fixed-size allocation per
// device, made on first use OUTSIDE stream capture and never freed or regrown: the pointers are baked into captured CUDA
// graphs, so they must stay valid for the life of the process.  The counters are zero when allocated and the kernel
// leaves them at zero (the last arriver of every counter resets it), so no memset sits between launches.  The library
// is single-stream per device: two fused launches never overlap.
static constexpr int kFusedFlagInts = 1 << 18;       // 2*row_tiles + 16*B ints per launch: up to ~130 k row tiles
static constexpr int kFusedPartials = 1 << 20;       // row_tiles * 8 floats per launch
struct FusedScratch {
  int* flags = nullptr;
  float* partials = nullptr;
};
inline FusedScratch& fused_scratch(cudaStream_t st) {
  static FusedScratch per_dev[64];
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) throw std::runtime_error("xattn_fused: device index out of range");
  FusedScratch& s = per_dev[dev];
  if (!s.flags) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    B200_CHECK(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone)
      throw std::runtime_error("xattn_fused: first use inside stream capture - run the launch sequence once eagerly "
                               "before capturing it");
    B200_CHECK(cudaMalloc(&s.flags, sizeof(int) * kFusedFlagInts));
    B200_CHECK(cudaMemset(s.flags, 0, sizeof(int) * kFusedFlagInts));
    B200_CHECK(cudaMalloc(&s.partials, sizeof(float) * kFusedPartials));
  }
  return s;
}
/* after a kernel fault (trap) the counters may be left non-zero: re-zero them (host-synchronous) */
inline void fused_scratch_reset() {
  FusedScratch& s = fused_scratch((cudaStream_t)0);
  B200_CHECK(cudaMemset(s.flags, 0, sizeof(int) * kFusedFlagInts));
}
inline CUtensorMap rowmajor_map_2d(const __half* p, long long rows, int cols, int ld, int box_rows) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t st[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {64, (uint32_t)box_rows};
  return make_tmap_f16(p, 2, dims, st, box);
}

template <int D>
inline void launch_fused_t(const CUtensorMap& tmX, const CUtensorMap& tmWq, const CUtensorMap& tmK,
                           const CUtensorMap& tmVt, const CUtensorMap& tmO, const CUtensorMap& tmWo,
                           const FusedXattnParams& p, int row_tiles, cudaStream_t st) {
  using Cfg = FusedCfg<D>;
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(xattn_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES));
    done = true;
  }
  // forward-progress guard for the cross-CTA spin-waits: the two clusters of a row tile (8 heads) and, with the loss on,
  // every row tile of an image must be able to be resident together.  Queried once per device (first launch is eager).
  static int max_clusters = -1;
  if (max_clusters < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1024);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 4; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    B200_CHECK(cudaOccupancyMaxActiveClusters(&n, xattn_fused_kernel<D>, &cfg));
    max_clusters = n;
  }
  const int need = p.has_loss ? 2 * p.tiles_per_img : 2;
  if (need > max_clusters)
    throw std::runtime_error("xattn_fused: the hand-shake needs " + std::to_string(need) + " co-resident clusters, the "
                             "device can hold " + std::to_string(max_clusters) + " (MIG / MPS slice?) - use the unfused path");
  xattn_fused_kernel<D><<<dim3(row_tiles * 8), 192, Cfg::SMEM_BYTES, st>>>(tmX, tmWq, tmK, tmVt, tmO, tmWo, p);
  B200_CHECK(cudaGetLastError());
}
}  // namespace b200

extern "C" int b200lmd_xattn_fused_supported(int heads, int head_dim, int n) {
  return heads == 8 && n % 128 == 0 && (head_dim == 64 || head_dim == 80 || head_dim == 160);
}

extern "C" int b200lmd_xattn_fused_f16(const void* x, const void* wq, const void* k_slab, const void* vt_slab,
                                       const void* wo, const void* bias_o, const void* residual, void* out,
                                       void* o_scratch, void* q_slab, void* lse2, void* probs, const int* save_tok,
                                       void* probs_tok, const b200lmd_xattn_loss* loss, int B, int n, int heads,
                                       int head_dim, int nk, int k_alloc, float scale, void* stream) {
  return b200::guarded([&] {
    using namespace b200;
    if (!b200lmd_xattn_fused_supported(heads, head_dim, n)) throw std::runtime_error("xattn_fused: unsupported shape");
    if (nk > 80 || k_alloc < 80) throw std::runtime_error("xattn_fused: needs <= 80 text keys in 80-row slabs");
    {
      if (loss && n > 1280) throw std::runtime_error("xattn_fused loss: at most 1280 query tokens per image");
    }
    const int C = heads * head_dim;
    const long long M = (long long)B * n;
    const int dp = round_dp(head_dim), d16 = round_d16(head_dim);
    if (d16 != head_dim) throw std::runtime_error("xattn_fused: head_dim must be a multiple of 16");
    FusedXattnParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.d = head_dim; p.C = C; p.nk = nk; p.k_alloc = k_alloc; p.tiles_per_img = n / 128;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.bias_o = (const float*)bias_o; p.residual = (const __half*)residual; p.out = (__half*)out;
    p.o_buf = (__half*)o_scratch; p.q_slab = (__half*)q_slab; p.lse2 = (float*)lse2; p.probs = (__half*)probs;
    p.save_tok = save_tok; p.probs_tok = (__half*)probs_tok;
    p.has_loss = loss != nullptr;
    if (loss) p.L = *reinterpret_cast<const XattnLoss*>(loss);
    p.dbg = fused_dbg();
    p.loss_stage = fused_loss_stage() ? 1 : 0;
    // per-row-tile arrival counters live right behind the scratch rows of o_scratch's owner: a small static buffer
    const int n_tiles = (int)(M / 128);
    if (2 * n_tiles + 16 * B > kFusedFlagInts || n_tiles * 8 > kFusedPartials)
      throw std::runtime_error("xattn_fused: batch exceeds the hand-shake scratch (split the launch)");
    FusedScratch& fs = fused_scratch((cudaStream_t)stream);
    p.tile_flags = fs.flags;
    p.tile_done = p.tile_flags + n_tiles;
    p.bh_ready = p.tile_done + n_tiles;
    p.bh_done = p.bh_ready + B * 8;
    p.loss_partials = fs.partials;
    CUtensorMap tmX = rowmajor_map_2d((const __half*)x, M, C, C, 32);
    CUtensorMap tmO = rowmajor_map_2d((const __half*)o_scratch, M, C, C, 32);
    CUtensorMap tmWq = rowmajor_map_2d((const __half*)wq, C, C, C, head_dim);
    CUtensorMap tmWo = rowmajor_map_2d((const __half*)wo, C, C, C, head_dim);
    uint64_t kd[3] = {(uint64_t)dp, (uint64_t)k_alloc, (uint64_t)B * heads};
    uint64_t ks[2] = {(uint64_t)dp * 2, (uint64_t)dp * 2 * k_alloc};
    uint32_t kb[3] = {64, 80, 1};
    CUtensorMap tmK = make_tmap_f16(k_slab, 3, kd, ks, kb);
    CUtensorMap tmVt = slab_tr_map((const __half*)vt_slab, B * heads, d16, k_alloc);
    const int row_tiles = (int)(M / 128);
    cudaStream_t st = (cudaStream_t)stream;
    switch (head_dim) {
      case 64: launch_fused_t<64>(tmX, tmWq, tmK, tmVt, tmO, tmWo, p, row_tiles, st); break;
      case 80: launch_fused_t<80>(tmX, tmWq, tmK, tmVt, tmO, tmWo, p, row_tiles, st); break;
      default: launch_fused_t<160>(tmX, tmWq, tmK, tmVt, tmO, tmWo, p, row_tiles, st); break;
    }
  });
}

/* re-zero the fused kernel's hand-shake counters after a faulted launch (host-synchronous) */
extern "C" int b200lmd_xattn_fused_reset(void) {
  return b200::guarded([&] { b200::fused_scratch_reset(); });
}

/* profiling aid: an EMPTY kernel with xattn_fused_kernel's launch configuration (grid of `ctas` CTAs in clusters of 4,
 * 192 threads, the same dynamic shared memory) - its event-bracketed duration is the launch + drain floor that the fused
 * kernel's own duration contains (floor model) */
namespace b200 {
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(192, 1) xattn_fused_null_kernel(int* sink) {
  if (sink && threadIdx.x == 0 && blockIdx.x == 0xFFFFFFu) *sink = 1;
}
}  // namespace b200
extern "C" int b200lmd_xattn_fused_launch_floor(int head_dim, int ctas, void* stream) {
  return b200::guarded([&] {
    using namespace b200;
    if (ctas <= 0 || ctas % 4) throw std::runtime_error("launch_floor: ctas must be a positive multiple of 4");
    const int smem = head_dim == 64 ? FusedCfg<64>::SMEM_BYTES : head_dim == 80 ? FusedCfg<80>::SMEM_BYTES
                                                                               : FusedCfg<160>::SMEM_BYTES;
    static bool done = false;
    if (!done) {
      B200_CHECK(cudaFuncSetAttribute(xattn_fused_null_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      FusedCfg<160>::SMEM_BYTES > FusedCfg<64>::SMEM_BYTES ? FusedCfg<160>::SMEM_BYTES
                                                                                         : FusedCfg<64>::SMEM_BYTES));
      done = true;
    }
    xattn_fused_null_kernel<<<dim3(ctas), 192, smem, (cudaStream_t)stream>>>(nullptr);
    B200_CHECK(cudaGetLastError());
  });
}

/* profiling aid: device buffer [grid][8] of %globaltimer stamps written by xattn_fused_kernel (NULL disables) */
extern "C" int b200lmd_set_debug_buffer(void* p) {
  b200::fused_dbg() = (unsigned long long*)p;
  return 0;
}
