// Host-side builders for gemm_tc_kernel launches: tensor-map encoding, brick selection, tap tables.
#pragma once
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "gemm.cuh"
#include "gemm2.cuh"
#include "gemm3.cuh"

namespace b200 {

#define B200_CHECK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));     \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p) throw std::runtime_error("cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 tensor map, 128B swizzle, zero OOB fill. dims/box innermost-first; strides in BYTES for dims 1..rank-1.
inline CUtensorMap make_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) throw std::runtime_error("tensor map base not 16B aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) throw std::runtime_error("tensor map stride not a multiple of 16 bytes");
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs,
                               bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu box=%u,%u", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    throw std::runtime_error(buf);
  }
  return m;
}

struct GemmLaunch {
  CUtensorMap tmA, tmB;
  GemmParams p;
  int block_n;
  dim3 grid;
};

inline int pow2_floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

// pick a (tb, th, tw) brick with tb*th*tw == 128 that minimises padded work
inline void pick_brick(int B, int H, int W, int& tb, int& th, int& tw) {
  long long best = -1;
  for (int cw = 128; cw >= 1; cw >>= 1) {
    for (int ch = 128 / cw; ch >= 1; ch >>= 1) {
      int cb = 128 / (cw * ch);
      long long padded = (long long)((W + cw - 1) / cw) * cw * ((H + ch - 1) / ch) * ch * ((B + cb - 1) / cb) * cb;
      if (best < 0 || padded < best) {
        best = padded;
        tb = cb; th = ch; tw = cw;
      }
    }
  }
}

// A: NHWC fp16 activations [B, H, W, ldc] using channels [c_off, c_off+Cin) ; Wt: [N][wtaps][Cin] fp16.
// The output grid (OB x OHl x OWl) is the set of pixels the M tiles enumerate; input pixel = output pixel + tap offset
// in A's own coordinate system (stride-2 patterns are expressed by pointing A at a space-to-depth copy).
struct GemmBuild {
  const __half* A = nullptr;
  int aB = 1, aH = 1, aW = 1, a_ld = 0, Cin = 0;  // A tensor geometry (a_ld = channel stride in elements)
  const __half* Wt = nullptr;
  int N = 0, wtaps = 1;                           // weight taps dimension (size of the middle dim)
  int gB = 1, gH = 1, gW = 1;                     // output-pixel grid walked by M tiles
  int ntaps = 1;
  GemmTap taps[9] = {};
  int force_block_n = 0;
};

inline GemmLaunch build_gemm(const GemmBuild& b, GemmParams ep /* epilogue fields prefilled */) {
  GemmLaunch L;
  memset(&L.tmA, 0, sizeof(L.tmA));
  memset(&L.tmB, 0, sizeof(L.tmB));
  GemmParams& p = ep;
  p.B = b.gB; p.H = b.gH; p.W = b.gW;
  pick_brick(b.gB, b.gH, b.gW, p.tb, p.th, p.tw);
  p.tiles_x = (b.gW + p.tw - 1) / p.tw;
  p.tiles_y = (b.gH + p.th - 1) / p.th;
  p.tiles_b = (b.gB + p.tb - 1) / p.tb;
  p.N = b.N;
  p.Cin = b.Cin;
  p.ntaps = b.ntaps;
  for (int i = 0; i < 9; ++i) p.taps[i] = b.taps[i];
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  int bn = b.force_block_n;
  if (!bn) {
    if (p.mode == EPI_GEGLU) bn = 128;
    else if (b.N <= 64) bn = 64;
    else {
      // 256-wide tiles halve the A-operand traffic per FLOP (the short-K projections are bound by bytes in flight from
      // L2, not by the tensor pipe); taken when the padded last tile wastes at most a tenth of the columns and the
      // grid still fills the machine
      const int n256 = (b.N + 255) / 256;
      const bool small_waste = (n256 * 256 - b.N) * 10 <= b.N;
      bn = (b.N > 256 && small_waste && (long long)m_tiles * n256 >= 120) ? 256 : 128;
    }
  }
  L.block_n = bn;
  {
    uint64_t dims[4] = {(uint64_t)b.Cin, (uint64_t)b.aW, (uint64_t)b.aH, (uint64_t)b.aB};
    uint64_t strides[3] = {(uint64_t)b.a_ld * 2, (uint64_t)b.a_ld * 2 * b.aW, (uint64_t)b.a_ld * 2 * b.aW * b.aH};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
    L.tmA = make_tmap_f16(b.A, 4, dims, strides, box);
  }
  {
    uint64_t dims[3] = {(uint64_t)b.Cin, (uint64_t)b.wtaps, (uint64_t)b.N};
    uint64_t strides[2] = {(uint64_t)b.Cin * 2, (uint64_t)b.Cin * 2 * b.wtaps};
    uint32_t box[3] = {64, 1, (uint32_t)bn};
    L.tmB = make_tmap_f16(b.Wt, 3, dims, strides, box);
  }
  L.p = p;
  L.grid = dim3((b.N + bn - 1) / bn, m_tiles, 1);
  return L;
}

template <int BN>
inline void gemm_set_attr() {
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    GemmCfg<BN>::SMEM_BYTES));
    done = true;
  }
}

inline bool& gemm_use_v2() {
  static bool v = true;   // persistent kernel with the coalesced epilogue (gemm2.cuh)
  return v;
}

template <int BN, int MODE>
inline void launch_gemm2_mode(const GemmLaunch& L, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(gemm2_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Gemm2Cfg<BN>::SMEM_BYTES));
    done = true;
  }
  const int m_tiles = (int)L.grid.y, n_tiles = (int)L.grid.x;
  const long long total = (long long)m_tiles * n_tiles;
  const int grid = (int)(total < 148 ? total : 148);
  gemm2_tc_kernel<BN, MODE><<<grid, 320, Gemm2Cfg<BN>::SMEM_BYTES, st>>>(L.tmA, L.tmB, L.p, m_tiles, n_tiles);
  B200_CHECK(cudaGetLastError());
}
template <int BN>
inline void launch_gemm2(const GemmLaunch& L, cudaStream_t st) {
  switch (L.p.mode) {
    case EPI_ROWMAJOR: launch_gemm2_mode<BN, EPI_ROWMAJOR>(L, st); break;
    case EPI_GEGLU:
      if (BN != 128) throw std::runtime_error("gemm2: the GEGLU epilogue needs 128-wide tiles");
      launch_gemm2_mode<128, EPI_GEGLU>(L, st);
      break;
    case EPI_HEADS: launch_gemm2_mode<BN, EPI_HEADS>(L, st); break;
    default: throw std::runtime_error("gemm2: unknown epilogue mode");
  }
}

inline bool& gemm_use_v3() {
  static bool v = true;   // TMA-only epilogue (gemm3.cuh) for row-major fp16 outputs
  return v;
}

inline bool gemm3_eligible(const GemmLaunch& L) {
  const GemmParams& p = L.p;
  // measured (profiles/bench_kernels.py): the TMA epilogue wins on short contractions (K=320: 80 -> 56 us), while long
  // ones prefer the deeper smem pipeline of gemm2 (conv 3x3: 147 vs 153 us)
  const int num_kb = ((p.Cin + 63) / 64) * p.ntaps;
  if (num_kb > 10) return false;
  return gemm_use_v3() && L.block_n >= 128 && p.mode == EPI_ROWMAJOR && p.out && !p.out_f32 && (p.N % 8 == 0) &&
         (p.ldo % 8 == 0) && p.sy == 1 && p.sx == 1 && p.oy == 0 && p.ox == 0 && p.OH == p.H && p.OW == p.W &&
         (!p.residual || p.ldr % 8 == 0) && !(p.residual && p.accumulate_out);
}

// 4-D view (N, W, H, B) of a row-major [rows, ld] tensor whose rows enumerate the output-pixel grid
inline CUtensorMap out_tile_map(const __half* base, int ld, const GemmParams& p) {
  uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * p.W, (uint64_t)ld * 2 * p.W * p.H};
  uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tb};
  return make_tmap_f16(base, 4, dims, strides, box);
}

template <int BN>
inline void launch_gemm3(const GemmLaunch& L, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    B200_CHECK(cudaFuncSetAttribute(gemm3_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Gemm3Cfg<BN>::SMEM_BYTES));
    done = true;
  }
  const GemmParams& p = L.p;
  const CUtensorMap tmC = out_tile_map(p.out, p.ldo, p);
  const CUtensorMap tmR = p.residual ? out_tile_map(p.residual, p.ldr, p) : tmC;   // accumulate: re-read the output
  const int m_tiles = (int)L.grid.y, n_tiles = (int)L.grid.x;
  const long long total = (long long)m_tiles * n_tiles;
  const int grid = (int)(total < 148 ? total : 148);
  gemm3_tc_kernel<BN><<<grid, 320, Gemm3Cfg<BN>::SMEM_BYTES, st>>>(L.tmA, L.tmB, tmC, tmR, L.p, m_tiles, n_tiles);
  B200_CHECK(cudaGetLastError());
}

inline bool gemm2_eligible(const GemmLaunch& L) {
  const GemmParams& p = L.p;
  if (!gemm_use_v2() || L.block_n < 128 || p.out_f32 || (p.N % 8)) return false;
  if (p.mode == EPI_ROWMAJOR) return p.out && (p.ldo % 8 == 0) && (!p.residual || p.ldr % 8 == 0);
  if (p.mode == EPI_GEGLU) return L.block_n == 128;
  return true;  // EPI_HEADS
}

inline void run_gemm(const GemmLaunch& L, cudaStream_t st) {
  if (gemm3_eligible(L)) {
    if (L.block_n == 128) launch_gemm3<128>(L, st);
    else launch_gemm3<256>(L, st);
    return;
  }
  if (gemm2_eligible(L)) {
    if (L.block_n == 128) launch_gemm2<128>(L, st);
    else launch_gemm2<256>(L, st);
    return;
  }
  switch (L.block_n) {
    case 64:
      gemm_set_attr<64>();
      gemm_tc_kernel<64><<<L.grid, 192, GemmCfg<64>::SMEM_BYTES, st>>>(L.tmA, L.tmB, L.p);
      break;
    case 128:
      gemm_set_attr<128>();
      gemm_tc_kernel<128><<<L.grid, 192, GemmCfg<128>::SMEM_BYTES, st>>>(L.tmA, L.tmB, L.p);
      break;
    case 256:
      gemm_set_attr<256>();
      gemm_tc_kernel<256><<<L.grid, 192, GemmCfg<256>::SMEM_BYTES, st>>>(L.tmA, L.tmB, L.p);
      break;
    default:
      throw std::runtime_error("bad block_n");
  }
  B200_CHECK(cudaGetLastError());
}

inline GemmParams default_epilogue() {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.mode = EPI_ROWMAJOR;
  p.alpha = 1.f;
  p.sy = p.sx = 1;
  return p;
}

// 3x3 / pad 1 / stride 1 taps, weight tap index = r*3+s
inline void fill_taps_3x3(GemmBuild& b) {
  b.ntaps = 9;
  b.wtaps = 9;
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) b.taps[r * 3 + s] = GemmTap{(int16_t)(s - 1), (int16_t)(r - 1), 0, (int16_t)(r * 3 + s)};
}

}  // namespace b200
