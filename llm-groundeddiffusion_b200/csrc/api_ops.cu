// C-ABI entry points for the stand-alone dense ops (see include/b200lmd.h).
#include <string>

#include "../../include/b200lmd.h"
#include "attention_host.cuh"
#include "boxdiff.cuh"
#include "elementwise.cuh"
#include "gemm_host.cuh"

namespace b200 {
thread_local std::string g_last_error;

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  } catch (...) {
    g_last_error = "unknown error";
    return 2;
  }
}

__global__ void geglu_interleave_w_kernel(const __half* __restrict__ w, __half* __restrict__ o, int F, int K) {
  // output row t*128 + j (j<64)  <- value row t*64+j ; t*128+64+j <- gate row F + t*64 + j
  const long long total = 2LL * F * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long orow = i / K;
    const int k = (int)(i % K);
    const int t = (int)(orow / 128), j = (int)(orow % 128);
    const long long src = (j < 64) ? (long long)(t * 64 + j) : (long long)F + t * 64 + (j - 64);
    o[i] = w[src * K + k];
  }
}
__global__ void geglu_interleave_b_kernel(const float* __restrict__ b, float* __restrict__ o, int F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * F) {
    const int t = i / 128, j = i % 128;
    o[i] = b[(j < 64) ? (t * 64 + j) : (F + t * 64 + j - 64)];
  }
}
}  // namespace b200

using namespace b200;

extern "C" const char* b200lmd_last_error(void) { return g_last_error.c_str(); }
extern "C" int b200lmd_version(void) { return 100; }

extern "C" int b200lmd_linear_f16(const void* x, int ldx, const void* w, const void* bias, const void* residual,
                                  int ldr, void* y, int ldy, void* y_f32, int M, int N, int K, float alpha,
                                  int accumulate, void* stream) {
  return guarded([&] {
    GemmBuild b;
    b.A = (const __half*)x; b.aB = 1; b.aH = 1; b.aW = M; b.a_ld = ldx; b.Cin = K;
    b.Wt = (const __half*)w; b.N = N; b.wtaps = 1; b.ntaps = 1;
    b.gB = 1; b.gH = 1; b.gW = M;
    b.taps[0] = GemmTap{0, 0, 0, 0};
    GemmParams ep = default_epilogue();
    ep.alpha = alpha;
    ep.bias = (const float*)bias;
    ep.residual = (const __half*)residual; ep.ldr = ldr;
    ep.out = (__half*)y; ep.ldo = ldy;
    ep.out_f32 = (float*)y_f32; ep.ldo32 = ldy;
    ep.accumulate_out = accumulate;
    ep.OH = 1; ep.OW = M;
    run_gemm(build_gemm(b, ep), (cudaStream_t)stream);
  });
}

extern "C" int b200lmd_geglu_interleave_w(const void* w, void* w_il, int F, int K, void* stream) {
  return guarded([&] {
    if (F % 64) throw std::runtime_error("GEGLU inner dim must be a multiple of 64");
    geglu_interleave_w_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>((const __half*)w, (__half*)w_il, F, K);
    B200_CHECK(cudaGetLastError());
  });
}
extern "C" int b200lmd_geglu_interleave_b(const void* bias, void* bias_il, int F, void* stream) {
  return guarded([&] {
    geglu_interleave_b_kernel<<<(2 * F + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float*)bias, (float*)bias_il, F);
    B200_CHECK(cudaGetLastError());
  });
}

extern "C" int b200lmd_linear_geglu_f16(const void* x, int ldx, const void* w_il, const void* bias_il, void* y,
                                        void* pre, int M, int F, int K, void* stream) {
  return guarded([&] {
    GemmBuild b;
    b.A = (const __half*)x; b.aW = M; b.a_ld = ldx; b.Cin = K;
    b.Wt = (const __half*)w_il; b.N = 2 * F;
    b.gW = M;
    b.taps[0] = GemmTap{0, 0, 0, 0};
    GemmParams ep = default_epilogue();
    ep.mode = EPI_GEGLU;
    ep.bias = (const float*)bias_il;
    ep.out = (__half*)y; ep.ldo = F;
    ep.pre = (__half*)pre;
    ep.OH = 1; ep.OW = M;
    run_gemm(build_gemm(b, ep), (cudaStream_t)stream);
  });
}

extern "C" int b200lmd_conv3x3_f16(const void* x, const void* w, const void* bias, const void* chan_add,
                                   const void* residual, void* y, void* y_f32, int B, int H, int W, int Cin, int Cout,
                                   void* stream) {
  return guarded([&] {
    GemmBuild b;
    b.A = (const __half*)x; b.aB = B; b.aH = H; b.aW = W; b.a_ld = Cin; b.Cin = Cin;
    b.Wt = (const __half*)w; b.N = Cout;
    b.gB = B; b.gH = H; b.gW = W;
    fill_taps_3x3(b);
    GemmParams ep = default_epilogue();
    ep.bias = (const float*)bias;
    ep.chan_add = (const float*)chan_add;
    ep.rows_per_img = H * W;
    ep.residual = (const __half*)residual; ep.ldr = Cout;
    ep.out = (__half*)y; ep.ldo = Cout;
    ep.out_f32 = (float*)y_f32; ep.ldo32 = Cout;
    ep.OH = H; ep.OW = W;
    run_gemm(build_gemm(b, ep), (cudaStream_t)stream);
  });
}

extern "C" int b200lmd_round_dp(int d) { return round_dp(d); }
extern "C" int b200lmd_round_d16(int d) {
  try { return round_d16(d); } catch (...) { return -1; }
}

extern "C" int b200lmd_project_heads2_f16(const void* x, int ldx, const void* w, int M, int N, int K,
                                          int rows_per_img, int heads, int head_dim, int which0, void* const* rm,
                                          const int* rm_alloc, void* const* tr, const int* tr_alloc, void* stream) {
  return guarded([&] {
    if (head_dim % 8) throw std::runtime_error("head_dim must be a multiple of 8");
    GemmBuild b;
    b.A = (const __half*)x; b.aW = M; b.a_ld = ldx; b.Cin = K;
    b.Wt = (const __half*)w; b.N = N;
    b.gW = M;
    b.taps[0] = GemmTap{0, 0, 0, 0};
    GemmParams ep = default_epilogue();
    ep.mode = EPI_HEADS;
    ep.rows_per_img = rows_per_img;
    ep.C = heads * head_dim; ep.heads = heads; ep.d = head_dim; ep.which0 = which0;
    ep.dp = round_dp(head_dim); ep.d16 = round_d16(head_dim);
    for (int i = 0; i < 3; ++i) {
      ep.rm[i] = (__half*)rm[i]; ep.rm_alloc[i] = rm_alloc[i];
      ep.tr[i] = (__half*)tr[i]; ep.tr_alloc[i] = tr_alloc[i];
    }
    ep.OH = 1; ep.OW = M;
    run_gemm(build_gemm(b, ep), (cudaStream_t)stream);
  });
}

extern "C" int b200lmd_attention_bwd_f16(const void* q, const void* k, const void* v, const void* dO, const void* qt,
                                         const void* kt, const void* dOt, const void* lse2, void* delta,
                                         const void* o_tok, int ld_o, const void* do_tok, int ld_do,
                                         const void* dp_extra, int ext_ld, void* dq, int ld_dq, void* dk, int ld_dk,
                                         void* dv, int ld_dv, int nk_store, int B, int heads, int nq, int nk,
                                         int q_alloc, int k_alloc, int head_dim, float scale, void* stream) {
  return guarded([&] {
    cudaStream_t st = (cudaStream_t)stream;
    const float* dl = nullptr;
    if (delta && !o_tok) dl = (const float*)delta;  // precomputed by the caller (b200lmd_attn_delta_slab)
    if (delta && o_tok) {
      run_attn_delta((const __half*)do_tok, ld_do, (const __half*)o_tok, ld_o, (float*)delta, B, heads, nq, q_alloc,
                     head_dim, st);
      dl = (const float*)delta;
    }
    run_attn_dq(build_attn_dq((const __half*)q, (const __half*)dO, (const __half*)k, (const __half*)v,
                              (const __half*)kt, (const float*)lse2, dl, (const float*)dp_extra, ext_ld, (__half*)dq,
                              ld_dq, B, heads, nq, nk, q_alloc, k_alloc, head_dim, scale),
                st);
    if (dk || dv) {
      if (!dl) throw std::runtime_error("dK/dV need the delta buffer");
      run_attn_dkv(build_attn_dkv((const __half*)k, (const __half*)v, (const __half*)q, (const __half*)dO,
                                  (const __half*)qt, (const __half*)dOt, (const float*)lse2, dl, (__half*)dk, ld_dk,
                                  (__half*)dv, ld_dv, nk_store, B, heads, nq, nk, q_alloc, k_alloc, head_dim, scale),
                   st);
    }
  });
}

extern "C" int b200lmd_project_heads_f16(const void* x, int ldx, const void* w, int M, int N, int K, int rows_per_img,
                                         int heads, int head_dim, int which0, void* q, int q_alloc, void* k,
                                         int k_alloc, void* vt, int v_alloc, void* stream) {
  return guarded([&] {
    if (head_dim % 8) throw std::runtime_error("head_dim must be a multiple of 8");
    GemmBuild b;
    b.A = (const __half*)x; b.aW = M; b.a_ld = ldx; b.Cin = K;
    b.Wt = (const __half*)w; b.N = N;
    b.gW = M;
    b.taps[0] = GemmTap{0, 0, 0, 0};
    GemmParams ep = default_epilogue();
    ep.mode = EPI_HEADS;
    ep.rows_per_img = rows_per_img;
    ep.C = heads * head_dim; ep.heads = heads; ep.d = head_dim; ep.which0 = which0;
    ep.dp = round_dp(head_dim); ep.d16 = round_d16(head_dim);
    ep.rm[0] = (__half*)q; ep.rm_alloc[0] = q_alloc;
    ep.rm[1] = (__half*)k; ep.rm_alloc[1] = k_alloc;
    ep.tr[2] = (__half*)vt; ep.tr_alloc[2] = v_alloc;
    ep.OH = 1; ep.OW = M;
    run_gemm(build_gemm(b, ep), (cudaStream_t)stream);
  });
}

extern "C" int b200lmd_attention_fwd_f16(const void* q, const void* k, const void* vt, void* out, int ldo, void* lse2,
                                         int B, int heads, int nq, int nk, int q_alloc, int k_alloc, int head_dim,
                                         float scale, void* stream) {
  return guarded([&] {
    run_attn_fwd(build_attn_fwd((const __half*)q, (const __half*)k, (const __half*)vt, (__half*)out, ldo,
                                (float*)lse2, B, heads, nq, nk, q_alloc, k_alloc, head_dim, scale),
                 (cudaStream_t)stream);
  });
}

static int gn_rows_per(int B, int n) {
  // aim for ~4 blocks per SM over (chunks x B): several 256-thread blocks are resident per SM and hide each other's
  // load latency
  int chunks = (4 * kNumSMs + B - 1) / B;
  int rp = (n + chunks - 1) / chunks;
  return rp < 4 ? 4 : rp;
}

// GroupNorm reduction scratch (per device, allocated once): block partials + one ticket counter per image.  The
// counters reset themselves at the end of every launch; launches on one stream are serial, which is the only use here.
constexpr int kGnPartFloats = 1 << 20;
constexpr int kGnCounters = 1 << 16;
struct GnScratch {
  float* part = nullptr;
  int* counter = nullptr;
};
static GnScratch& gn_scratch(cudaStream_t st) {
  static GnScratch per_dev[64];
  int dev = 0;
  B200_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) throw std::runtime_error("GroupNorm: device index out of range");
  GnScratch& s = per_dev[dev];
  if (!s.part) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    B200_CHECK(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone)
      throw std::runtime_error("GroupNorm: first use inside stream capture - run the launch sequence once eagerly "
                               "before capturing it");
    B200_CHECK(cudaMalloc(&s.part, sizeof(float) * kGnPartFloats));
    B200_CHECK(cudaMalloc(&s.counter, sizeof(int) * kGnCounters));
    B200_CHECK(cudaMemset(s.counter, 0, sizeof(int) * kGnCounters));
  }
  return s;
}
static size_t gn_reduce_smem(int C) {
  const int vecs = C >> 3, tpr = vecs < 256 ? vecs : 256, rif = 256 / tpr;
  size_t b = sizeof(float) * 2 * rif * C;
  return b < 1024 ? 1024 : b;
}
static void gn_check_scratch(int B, int chunks, int groups) {
  if ((long long)B * chunks * groups * 2 > kGnPartFloats || B > kGnCounters || groups * 8 > 256)
    throw std::runtime_error("GroupNorm: batch x chunks x groups exceeds the reduction scratch");
}

extern "C" int b200lmd_groupnorm_f16(const void* x, const void* gamma, const void* beta, void* y, void* sums, int B,
                                     int n, int C, int groups, float eps, int silu, void* stream) {
  return guarded([&] {
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 || C % groups) throw std::runtime_error("GroupNorm: C must be a multiple of 8 and of groups");
    const int rp = gn_rows_per(B, n);
    dim3 grid((n + rp - 1) / rp, B);
    GnScratch& gs = gn_scratch(st);
    gn_check_scratch(B, grid.x, groups);
    gn_stats_kernel<<<grid, 256, gn_reduce_smem(C), st>>>((const __half*)x, (float*)sums, gs.part, gs.counter, n, C,
                                                          groups, rp);
    if (2 * C * sizeof(float) > 48 * 1024) throw std::runtime_error("GroupNorm: C too large for the scale/shift table");
    gn_apply_kernel<<<grid, 256, 2 * C * sizeof(float), st>>>((const __half*)x, (const float*)sums,
                                                              (const float*)gamma, (const float*)beta, (__half*)y, B, n,
                                                              C, groups, eps, silu, rp);
    B200_CHECK(cudaGetLastError());
  });
}

extern "C" int b200lmd_groupnorm_bwd_f16(const void* dy, const void* x, const void* sums, const void* gamma,
                                         const void* beta, void* dx, void* bsums, int B, int n, int C, int groups,
                                         float eps, int silu, int accumulate, void* stream) {
  return guarded([&] {
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 || C % groups) throw std::runtime_error("GroupNorm: C must be a multiple of 8 and of groups");
    const int rp = gn_rows_per(B, n);
    dim3 grid((n + rp - 1) / rp, B);
    GnScratch& gs = gn_scratch(st);
    gn_check_scratch(B, grid.x, groups);
    gn_bwd_stats_kernel<<<grid, 256, gn_reduce_smem(C), st>>>((const __half*)dy, (const __half*)x, (const float*)sums,
                                                              (const float*)gamma, (const float*)beta, (float*)bsums,
                                                              gs.part, gs.counter, n, C, groups, eps, silu, rp);
    gn_bwd_apply_kernel<<<grid, 256, 0, st>>>(
        (const __half*)dy, (const __half*)x, (const float*)sums, (const float*)bsums, (const float*)gamma,
        (const float*)beta, (__half*)dx, B, n, C, groups, eps, silu, accumulate, rp);
    B200_CHECK(cudaGetLastError());
  });
}

extern "C" int b200lmd_layernorm_f16(const void* x, const void* gamma, const void* beta, void* y, void* stats,
                                     long long rows, int C, float eps, void* stream) {
  return guarded([&] {
    if (C % 8) throw std::runtime_error("LayerNorm: C must be a multiple of 8");
    const int vecs = C >> 3;
    cudaStream_t st = (cudaStream_t)stream;
    if (vecs <= 64)
      ln_fwd_rows_kernel<2, 4, 2><<<ew_grid((rows + 3) / 4 * 32, 256), 256, 0, st>>>(
          (const __half*)x, (const float*)gamma, (const float*)beta, (__half*)y, (float*)stats, rows, C, eps);
    else if (vecs <= 96)
      ln_fwd_rows_kernel<3, 2, 2><<<ew_grid((rows + 1) / 2 * 32, 256), 256, 0, st>>>(
          (const __half*)x, (const float*)gamma, (const float*)beta, (__half*)y, (float*)stats, rows, C, eps);
    else
      ln_fwd_kernel<<<ew_grid(rows * 32, 256), 256, 0, st>>>(
          (const __half*)x, (const float*)gamma, (const float*)beta, (__half*)y, (float*)stats, rows, C, eps);
    B200_CHECK(cudaGetLastError());
  });
}
extern "C" int b200lmd_layernorm_bwd_f16(const void* dy, const void* x, const void* stats, const void* gamma, void* dx,
                                         long long rows, int C, int accumulate, void* stream) {
  return guarded([&] {
    ln_bwd_kernel<<<ew_grid(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)dy, (const __half*)x, (const float*)stats, (const float*)gamma, (__half*)dx, rows, C, accumulate);
    B200_CHECK(cudaGetLastError());
  });
}
extern "C" int b200lmd_geglu_bwd_f16(const void* pre, const void* dy, void* dpre, long long rows, int F, void* stream) {
  return guarded([&] {
    geglu_bwd_kernel<<<ew_grid(rows * (F / 8)), 256, 0, (cudaStream_t)stream>>>((const __half*)pre, (const __half*)dy,
                                                                              (__half*)dpre, rows, F);
    B200_CHECK(cudaGetLastError());
  });
}
#include "api_xattn.cuh"
#include "api_more.cuh"

// ---------------------------------------------------------------------------------------------- BoxDiff loss
static_assert(sizeof(b200lmd_boxdiff) == sizeof(b200::BoxdiffParams), "C ABI struct mismatch");
static_assert(sizeof(b200lmd_boxdiff_term) == sizeof(b200::BoxdiffTerm), "C ABI struct mismatch");

extern "C" int b200lmd_boxdiff_loss(const b200lmd_boxdiff* bp, int B, void* stream) {
  return guarded([&] {
    const BoxdiffParams& p = *reinterpret_cast<const BoxdiffParams*>(bp);
    if (p.n_keys < 1 || p.n_keys > kBoxdiffMaxKeys) throw std::runtime_error("boxdiff: 1..8 guidance keys");
    if (p.side * p.side != p.n || p.n > 1024 || (p.n & 31)) throw std::runtime_error("boxdiff: n must be a square multiple of 32, <= 1024");
    if (p.T < 3 || p.T > p.ext_ld) throw std::runtime_error("boxdiff: token count out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const long long per_img = (long long)p.n * p.T;
    boxdiff_mean_kernel<<<dim3((unsigned)((per_img + 255) / 256), (unsigned)B), 256, 0, st>>>(p);
    B200_CHECK(cudaGetLastError());
    const size_t smem = sizeof(float) * (2 * p.n + 2 * p.side + 32) + sizeof(int) * (32 + 2 * p.side);
    boxdiff_loss_kernel<<<B, p.n, smem, st>>>(p);
    B200_CHECK(cudaGetLastError());
    boxdiff_scatter_kernel<<<ew_grid((long long)B * p.heads * per_img), 256, 0, st>>>(p, B);
    B200_CHECK(cudaGetLastError());
  });
}
