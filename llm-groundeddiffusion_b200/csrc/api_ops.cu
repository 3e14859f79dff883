// C-ABI entry points for the stand-alone dense ops (see include/b200lmd.h).
#include <string>

#include "../../include/b200lmd.h"
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
