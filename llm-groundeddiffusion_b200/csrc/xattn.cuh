// Fused cross-attention + layout-guidance loss (forward), sm_100a.
//
// One CTA = 128 query rows of one (image, head).  The text has T <= 128 keys, so the whole row of scores is one tcgen05
// tile: S = Q K^T (TMEM) -> softmax by one thread per row -> fp16 P to shared memory (A operand) -> O = P V (TMEM) ->
// O to HBM.  In the same kernel:
//   * requested attention maps leave as fp16 (full [BH, n, T] and/or one token column per image) - the reference's
//     save_attn_to_dict / return_token_ca_only contract (models/attention_processor.py:463-482);
//   * the columns of P that the guidance loss reads go to a small fp32 scratch; the LAST CTA of each (image, head)
//     to arrive (atomic ticket) evaluates the per-phrase top-k foreground/background energies and the
//     reference-attention L1 terms (utils/guidance.py:91-242) with warp/CTA reductions, writes the loss partial and
//     d(loss)/dP (consumed as dP_extra by attn_bwd_dq_kernel), so neither the maps nor an autograd graph exist.
// Loss tables are built on the host (llm-groundeddiffusion_b200/guidance.py) and restate scale_proportion / k_fg /
// k_bg / normalisers exactly (integers must match the reference bit for bit).
#pragma once
#include "ptx.cuh"

namespace b200 {

static constexpr int kMaxSlots = 48;
static constexpr int kMaxTerms = 80;     // loss terms per image (48 token slots + reference terms)

struct LossTerm {
  int type;     // 0 = energy (fg/bg top-k), 1 = reference-attention L1, 2 = ratio-based energy (w_fg carries the weight)
  int slot;     // which saved column of P
  int mask;     // mask id (row of `masks`)
  int k_fg, k_bg;
  float w_fg, w_bg;   // already include loss_scale and all normalisers (per head)
  float w_ref;
  int ref;      // ref-map id (row block of `refs`)
};

struct XattnLoss {
  const int* img_term_off;   // [B+1]
  const LossTerm* terms;
  const uint8_t* masks;      // [n_masks][n]
  const float* refs;         // [n_refs][heads][n]
  const int* slot_tok;       // [B][kMaxSlots], -1 = unused
  float* pcol;               // [BH][kMaxSlots][n]
  int* counters;             // [BH], zero on entry, self-resetting
  float* loss_part;          // [BH]
  float* dp_extra;           // [BH][n][ext_ld]
  int ext_ld;
  float gscale;
  float eps;
};

struct XattnParams {
  int heads, nq, nk;
  int nq_alloc, nk_alloc;
  int d;
  float scale_log2;
  __half* out; int ldo;
  float* lse2;               // [BH, nq_alloc] or null
  __half* probs;             // [BH, nq, nk] or null
  const int* save_tok;       // [B] token index per image (or null); column goes to probs_tok
  __half* probs_tok;         // [BH, nq]
  int has_loss;
  XattnLoss L;
};

template <int DPB, int D16>
struct XattnCfg {
  static constexpr int Q_BYTES = DPB * 16384;
  static constexpr int V_ATOM = D16 * 128;
  static constexpr int P_BYTES = 2 * 16384;
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + 2 * V_ATOM + P_BYTES + 1024 + 256;
};

__device__ __forceinline__ float block128_sum(float v, float* red, int tid) {
  // reduction over the 128 softmax threads (4 warps); red: >= 4 floats of shared scratch
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if ((tid & 31) == 0) red[tid >> 5] = v;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  return red[0] + red[1] + red[2] + red[3];
}


// Loss / gradient of one (image, head) from the published P columns (utils/guidance.py:91-242).  Called by the 128
// softmax threads (4 warps) of the last CTA of that (image, head); scratch: >= 24*n + 640 + 36*kMaxTerms bytes of shared memory (n <= 900 with the callers' buffers).
//
// Work is split into independent "problems", one warp each (round-robin): an energy term contributes two top-k
// selections (foreground / background), a reference term one normalised-L1.  Top-k sum by bisection on the float bit
// pattern (values are probabilities >= 0, so the unsigned order is the numeric order): 31 counting passes find the k-th
// largest value x_k; sum = sum(v > x_k) + (k - count(v > x_k)) * x_k; the gradient goes to the elements above x_k plus
// the lowest-index ties (the oracle's tie rule).  No block-wide barrier is needed inside a problem.
// ---- loss reduction: one warp per problem ((phrase, side) energy or one reference-attention term) with the whole map
// column in registers.  Lane L owns the contiguous index block [L*per, L*per+per) so that "lowest index first" among
// ties is a prefix over lanes.
template <int PERMAX>
struct LossRaw {
  float v[PERMAX];     // P column values (fp16-representable)
  float r[PERMAX];     // reference map values (reference terms only)
  uint64_t m;          // bit j: mask byte of element j is non-zero
};

struct LossProb {      // decoded problem
  const float* src;
  const uint8_t* mk;
  const float* R;      // nullptr for energy problems
  int tok, side, k;    // side: 0 fg top-k, 1 bg top-k, 2 ratio-based energy
  float w;
  bool staged;         // src / mk point into shared memory (staged copy) instead of global memory
};

// staged copy of one (image, head)'s loss inputs in shared memory: the used pcol rows and one mask row per term
struct LossStaged {
  const float* col;    // [n_slots][n] or nullptr
  const uint8_t* mk;   // [n_terms][n]
};

// 8 mask bytes (one per element of a lane's block) -> bit j set when byte j is non-zero
__device__ __forceinline__ uint64_t loss_mask_bits8(uint64_t raw) {
  uint64_t m = 0ull;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if ((raw >> (8 * j)) & 0xFFull) m |= 1ull << j;
  return m;
}
// true when the lane's 8 values / 8 mask bytes of an energy problem can be fetched with three vector loads
__device__ __forceinline__ bool loss_vec_ok(const LossProb& P, int n, int per, int i0) {
  return per == 8 && i0 + 8 <= n && !P.R && !P.staged &&
         (((uintptr_t)(P.src + i0) & 15) | ((uintptr_t)(P.mk + i0) & 7)) == 0;
}

template <int PERMAX>
__device__ __forceinline__ void loss_load(LossRaw<PERMAX>& q, const LossProb& P, int n, int per, int lane) {
  const int i0 = lane * per;
  q.m = 0ull;
  if (PERMAX == 8 && loss_vec_ok(P, n, per, i0)) {
    // n = 256 (the guidance resolution of SD1.x at 512^2): the lane's 8 values and 8 mask bytes are three vector loads
    const float4 a = __ldcg(reinterpret_cast<const float4*>(P.src + i0));
    const float4 c = __ldcg(reinterpret_cast<const float4*>(P.src + i0) + 1);
    const uint2 mb = *reinterpret_cast<const uint2*>(P.mk + i0);
    q.v[0] = a.x; q.v[1] = a.y; q.v[2] = a.z; q.v[3] = a.w;
    q.v[4] = c.x; q.v[5] = c.y; q.v[6] = c.z; q.v[7] = c.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) q.r[j] = 0.f;
    q.m = loss_mask_bits8(((uint64_t)mb.y << 32) | mb.x);
    return;
  }
#pragma unroll
  for (int j = 0; j < PERMAX; ++j) {
    const int i = i0 + j;
    q.v[j] = 0.f;
    q.r[j] = 0.f;
    if (j < per && i < n) {
      const uint8_t mb = P.mk[i];
      q.v[j] = P.staged ? P.src[i] : __ldcg(P.src + i);
      if (P.R) q.r[j] = P.R[i];
      if (mb) q.m |= 1ull << j;
    }
  }
}

template <int PERMAX>
__device__ __forceinline__ float loss_energy(const LossRaw<PERMAX>& q, const LossProb& P, float* dpx, int ext_ld, int n,
                                             int per, int lane, float gscale) {
  const int i0 = lane * per;
  float* dpx_tok = dpx + P.tok;
  const int k = P.k;
  // entry: bits 0-15 fp16 pattern of the masked value (0 on the other side of the mask), bit 16 = on this side,
  // bit 17 = a real element (i < n)
  uint32_t e[PERMAX];
#pragma unroll
  for (int j = 0; j < PERMAX; ++j) {
    uint32_t x = 0u;
    if (j < per && i0 + j < n) {
      x = 0x20000u;
      if ((((q.m >> j) & 1ull) != 0ull) == (P.side == 0))
        x |= 0x10000u | (uint32_t)__half_as_ushort(__float2half_rn(q.v[j]));
    }
    e[j] = x;
  }
  // bisection for the k-th largest fp16 bit pattern (non-negative values: unsigned order == numeric order):
  // invariant count(v >= lo) >= k, count(v >= hi) < k; 15 counting passes over registers
  uint32_t lo = 0u, hi = 0x7C00u;
  while (hi - lo > 1u) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
#pragma unroll
    for (int j = 0; j < PERMAX; ++j) c += ((e[j] & 0xFFFFu) >= mid);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) lo = mid; else hi = mid;
  }
  const float xk = __half2float(__ushort_as_half((unsigned short)lo));
  int c_gt = 0, ties_here = 0;
  float s_gt = 0.f;
#pragma unroll
  for (int j = 0; j < PERMAX; ++j) {
    const uint32_t bits = e[j] & 0xFFFFu;
    if (bits > lo) { ++c_gt; s_gt += __half2float(__ushort_as_half((unsigned short)bits)); }
    ties_here += ((e[j] & 0x20000u) && bits == lo);
  }
  c_gt = __reduce_add_sync(0xffffffffu, c_gt);
#pragma unroll
  for (int o = 16; o; o >>= 1) s_gt += __shfl_xor_sync(0xffffffffu, s_gt, o);
  int excl = ties_here;                                    // exclusive prefix of tie counts over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, excl, o);
    if (lane >= o) excl += up;
  }
  excl -= ties_here;
  const int need = k - c_gt;                               // ties to take (>= 1)
  const float tot = s_gt + (float)need * xk;
  const float gval = (P.side ? P.w : -P.w) / (float)k * gscale;
  int tr = excl;
#pragma unroll
  for (int j = 0; j < PERMAX; ++j) {
    const uint32_t bits = e[j] & 0xFFFFu;
    bool sel = bits > lo;
    if ((e[j] & 0x20000u) && bits == lo) { sel = tr < need; ++tr; }
    if (sel && (e[j] & 0x10000u)) atomicAdd(dpx_tok + (long long)(i0 + j) * ext_ld, gval);
  }
  return P.side ? P.w * tot / (float)k : P.w * (1.f - tot / (float)k);
}

// reference-attention L1 term: masked, sum-normalised maps compared in L1
template <int PERMAX>
__device__ __forceinline__ float loss_ref(const LossRaw<PERMAX>& q, const LossProb& P, float* dpx, int ext_ld, int lane,
                                          int per, float eps, float gscale) {
  const int i0 = lane * per;
  float* dpx_tok = dpx + P.tok;
  float sa = 0.f, sr = 0.f;
#pragma unroll
  for (int j = 0; j < PERMAX; ++j)
    if ((q.m >> j) & 1ull) { sa += q.v[j]; sr += q.r[j]; }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
  }
  const float A = sa + eps, Rs = sr + eps;
  float l1 = 0.f, inner = 0.f;
#pragma unroll
  for (int j = 0; j < PERMAX; ++j)
    if ((q.m >> j) & 1ull) {
      const float ah = q.v[j] / A, df = ah - q.r[j] / Rs;
      const float sg = (df > 0.f) - (df < 0.f);
      l1 += fabsf(df);
      inner += sg * ah;
    }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    inner += __shfl_xor_sync(0xffffffffu, inner, o);
  }
#pragma unroll
  for (int j = 0; j < PERMAX; ++j)
    if ((q.m >> j) & 1ull) {
      const float df = q.v[j] / A - q.r[j] / Rs;
      const float sg = (df > 0.f) - (df < 0.f);
      atomicAdd(dpx_tok + (long long)(i0 + j) * ext_ld, P.w / A * (sg - inner) * gscale);
    }
  return P.w * l1;
}

// ratio-based energy (utils/guidance.py:122-128, the deprecated default that generation/backward_guidance.py still runs):
// a = sum_q P M / sum_q P per head, loss = w (1 - a)^2 (w holds scale / (heads T_o n_obj n_keys));
// d loss / d P_q = -2 w (1 - a) (M_q - a) / sum_q P  - dense over the map
template <int PERMAX>
__device__ __forceinline__ float loss_ratio(const LossRaw<PERMAX>& q, const LossProb& P, float* dpx, int ext_ld, int n,
                                            int per, int lane, float gscale) {
  const int i0 = lane * per;
  float* dpx_tok = dpx + P.tok;
  float sa = 0.f, ss = 0.f;
#pragma unroll
  for (int j = 0; j < PERMAX; ++j) {
    ss += q.v[j];
    if ((q.m >> j) & 1ull) sa += q.v[j];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  // a column whose fp16 probabilities all underflowed to zero carries no information: the reference's fp16 path
  // divides 0 by 0 there (NaN loss); here the term counts as a = 0 with zero gradient (stated deviations)
  const float a = ss > 0.f ? sa / ss : 0.f;
  const float g0 = ss > 0.f ? -2.f * P.w * (1.f - a) / ss * gscale : 0.f;
#pragma unroll
  for (int j = 0; j < PERMAX; ++j)
    if (j < per && i0 + j < n) {
      const float mq = ((q.m >> j) & 1ull) ? 1.f : 0.f;
      atomicAdd(dpx_tok + (long long)(i0 + j) * ext_ld, g0 * (mq - a));
    }
  return P.w * (1.f - a) * (1.f - a);
}

constexpr int kLossScratchBytes = 160 * 4 + 160 * 2 + kMaxSlots * 4 + 16 + kMaxTerms * (int)sizeof(LossTerm);

__device__ __forceinline__ LossProb loss_decode(const XattnLoss& L, const LossTerm* sterms, const int* stok,
                                                unsigned short code, int bh, int h, int heads, int n,
                                                const LossStaged& stg) {
  const LossTerm& T = sterms[code >> 1];
  const int sub = code & 1;
  LossProb P;
  P.staged = stg.col != nullptr;
  P.src = P.staged ? stg.col + (long long)T.slot * n : L.pcol + ((long long)bh * kMaxSlots + T.slot) * n;
  P.mk = P.staged ? stg.mk + (long long)(code >> 1) * n : L.masks + (long long)T.mask * n;
  P.tok = stok[T.slot];
  if (T.type == 0) {
    P.R = nullptr;
    P.side = sub;
    P.k = sub ? T.k_bg : T.k_fg;
    P.w = sub ? T.w_bg : T.w_fg;
  } else if (T.type == 2) {
    P.R = nullptr;
    P.side = 2;
    P.k = 1;
    P.w = T.w_fg;
  } else {
    P.R = L.refs + ((long long)T.ref * heads + h) * n;
    P.side = 0;
    P.k = 1;
    P.w = T.w_ref;
  }
  return P;
}

template <int PERMAX, bool PIPE>
__device__ __forceinline__ void loss_problem_loop(const XattnLoss& L, const LossTerm* sterms, const int* stok,
                                                  const unsigned short* pcode, float* prob_loss, float* dpx, int n_prob,
                                                  int first, int stride, int lane, int bh, int h, int heads, int n,
                                                  int per, const LossStaged& stg, const LossRaw<PERMAX>& pre,
                                                  bool have_pre) {
  if (first >= n_prob) return;
  LossProb P = loss_decode(L, sterms, stok, pcode[first], bh, h, heads, n, stg);
  LossRaw<PERMAX> cur;
  if (have_pre) {                                          // first problem's inputs already requested by the caller
#pragma unroll
    for (int j = 0; j < PERMAX; ++j) {
      cur.v[j] = pre.v[j];
      cur.r[j] = 0.f;
    }
    cur.m = loss_mask_bits8(pre.m);                        // loss_prefetch_first leaves the 8 raw mask bytes here
  } else {
    loss_load(cur, P, n, per, lane);
  }
  for (int pid = first; pid < n_prob; pid += stride) {
    const bool more = pid + stride < n_prob;
    LossProb Pn = P;
    LossRaw<PERMAX> nxt;
    if (PIPE && more) {                                    // next problem's loads fly during this problem's compute
      Pn = loss_decode(L, sterms, stok, pcode[pid + stride], bh, h, heads, n, stg);
      loss_load(nxt, Pn, n, per, lane);
    }
    const float contrib = P.R ? loss_ref(cur, P, dpx, L.ext_ld, lane, per, L.eps, L.gscale)
                          : (P.side == 2 ? loss_ratio(cur, P, dpx, L.ext_ld, n, per, lane, L.gscale)
                                         : loss_energy(cur, P, dpx, L.ext_ld, n, per, lane, L.gscale));
    if (lane == 0) prob_loss[pid] = contrib;
    if (more) {
      if (PIPE) {
        cur = nxt;
        P = Pn;
      } else {
        P = loss_decode(L, sterms, stok, pcode[pid + stride], bh, h, heads, n, stg);
        loss_load(cur, P, n, per, lane);
      }
    }
  }
}

// ---- loss reduction of one (image, head), three steps run by 128 threads (4 warps, named barrier 1):
//   loss_stage : term table, slot->token table and the problem list into shared memory (inputs only: can run early)
//   loss_zero  : zero rows of dp_extra (must be ordered before any CTA's atomics into them)
//   loss_run   : this CTA's share of the problems (part of nparts), partial sums combined in a fixed order
// scratch: kLossScratchBytes of shared memory (16-byte aligned).  n <= 1280, at most 160 problems per image.
struct LossScratch {
  float* prob_loss;            // [160] per-problem contributions
  unsigned short* pcode;       // [160] term*2 + side
  int* stok;                   // [kMaxSlots] slot -> token
  int* n_prob;                 // [4] (one used; keeps the table 16-byte aligned)
  LossTerm* sterms;
};
__device__ __forceinline__ LossScratch loss_scratch(float* scratch) {
  LossScratch S;
  S.prob_loss = scratch;
  S.pcode = reinterpret_cast<unsigned short*>(scratch + 160);
  S.stok = reinterpret_cast<int*>(S.pcode + 160);
  S.n_prob = S.stok + kMaxSlots;
  S.sterms = reinterpret_cast<LossTerm*>(S.n_prob + 4);
  return S;
}
static_assert(kLossScratchBytes + 16 <= 4096, "loss scratch");

__device__ __forceinline__ void loss_stage(const XattnLoss& L, float* scratch, int tid, int b) {
  const int lane = tid & 31;
  LossScratch S = loss_scratch(scratch);
  const int t0 = L.img_term_off[b], t1 = L.img_term_off[b + 1];
  const int nterms = min(t1 - t0, kMaxTerms);
  {
    const int* src = reinterpret_cast<const int*>(L.terms + t0);
    int* dst = reinterpret_cast<int*>(S.sterms);
    const int words = nterms * (int)(sizeof(LossTerm) / 4);
    for (int i = tid; i < words; i += 128) dst[i] = src[i];
    if (tid < kMaxSlots) S.stok[tid] = L.slot_tok[b * kMaxSlots + tid];
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  // problem list: energy term -> 2 consecutive ids (fg, bg), reference term -> 1 id.  Every warp builds the same list
  // (identical values to identical addresses).
  int n_prob = 0;
  for (int tb = 0; tb < nterms; tb += 32) {
    const int t = tb + lane;
    const int nsub = t < nterms ? ((S.sterms[t].type == 0) ? 2 : 1) : 0;
    int incl = nsub;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int first = n_prob + incl - nsub;
    for (int sub = 0; sub < nsub; ++sub)
      if (first + sub < 160) S.pcode[first + sub] = (unsigned short)(t * 2 + sub);
    n_prob += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (tid == 0) {
    S.n_prob[0] = min(n_prob, 160);
    S.n_prob[1] = nterms;
    int ns = 0;
    while (ns < kMaxSlots && S.stok[ns] >= 0) ++ns;
    S.n_prob[2] = ns;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ void loss_zero(const XattnLoss& L, int bh, int n, int row0, int rows, int tid) {
  float4* d4 = reinterpret_cast<float4*>(L.dp_extra + ((long long)bh * n + row0) * L.ext_ld);
  const int tot = (rows * L.ext_ld) >> 2;                // ext_ld is a multiple of 4, rows are 16-byte aligned
  for (int i = tid; i < tot; i += 128) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Request the inputs of this warp's FIRST problem (the loads loss_run would start with) - for callers that have
// something else to do between "the P columns are published" and loss_run (the fused kernel: a cluster barrier).
// Only issues loads: nothing here consumes a loaded value (out.m holds the 8 RAW mask bytes, decoded by the consumer),
// so the calling warp does not wait for the round trip.  Returns false when there is nothing to prefetch or the
// three-vector-load layout does not apply (reference terms, n != 256); then loss_run loads as usual.
__device__ __forceinline__ bool loss_prefetch_first(const XattnLoss& L, float* scratch, int tid, int h, int heads, int bh,
                                                    int n, int part, int nparts, LossRaw<8>& out) {
  const int warp = tid >> 5, lane = tid & 31;
  LossScratch S = loss_scratch(scratch);
  const int per = (n + 31) >> 5;
  const int first = part * 4 + warp;
  if (per > 8 || first >= S.n_prob[0]) return false;
  LossStaged stg;
  stg.col = nullptr;
  stg.mk = nullptr;
  const LossProb P = loss_decode(L, S.sterms, S.stok, S.pcode[first], bh, h, heads, n, stg);
  const int i0 = lane * per;
  if (!loss_vec_ok(P, n, per, i0)) return false;           // per-lane condition is warp-uniform: per, n, alignment
  const float4 a = __ldcg(reinterpret_cast<const float4*>(P.src + i0));
  const float4 c = __ldcg(reinterpret_cast<const float4*>(P.src + i0) + 1);
  const uint2 mb = *reinterpret_cast<const uint2*>(P.mk + i0);
  out.v[0] = a.x; out.v[1] = a.y; out.v[2] = a.z; out.v[3] = a.w;
  out.v[4] = c.x; out.v[5] = c.y; out.v[6] = c.z; out.v[7] = c.w;
  out.m = ((uint64_t)mb.y << 32) | mb.x;
  return true;
}

// second half of the two-part combine when loss_run was given ticket_out: the later of the two parts resets the counters
__device__ __forceinline__ void loss_ticket_reset(int ticket, int* done, int* ready, int bh) {
  // opaque consumer: written in C the comparison is hoisted up to the atomic and waits for its round trip there
  asm volatile(
      "{\n .reg .pred p;\n setp.eq.s32 p, %0, 1;\n @p st.global.u32 [%1], 0;\n @p st.global.u32 [%2], 0;\n}" ::"r"(ticket),
      "l"(done + bh), "l"(ready + bh)
      : "memory");
}

#define LOSS_STAMP(i) do { if (dbg && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[i] = t_; } } while (0)
// part/nparts: the problems are dealt round-robin to nparts CTAs x 4 warps.  With nparts > 1 every CTA leaves its
// partial in partials[bh*nparts + part] and the last one to arrive (done[bh]) adds them up in part order, so the sum
// does not depend on arrival order.  done[bh] must be zero on entry and is zero again on exit (so is ready[bh]).
__device__ __forceinline__ int loss_run(const XattnLoss& L, float* scratch, int tid, int h, int heads, int bh, int n,
                                         int part, int nparts, float* partials, int* done,
                                         unsigned long long* dbg, int* ready,
                                         uint8_t* stage, int stage_bytes, bool defer_ticket, const LossRaw<8>& pre,
                                         bool have_pre) {
  int ticket = -1;                                         // returned when defer_ticket: see loss_ticket_reset
  const int warp = tid >> 5, lane = tid & 31;
  LossScratch S = loss_scratch(scratch);
  const int n_prob = S.n_prob[0];
  float* dpx = L.dp_extra + (long long)bh * n * L.ext_ld;
  const int per = (n + 31) >> 5;
  LOSS_STAMP(0);
  // Stage the inputs of every problem of this (image, head) in shared memory with ONE cooperative pass of 16-byte
  // loads (the used pcol rows are contiguous; one mask row per term), so the per-problem loads hit shared memory
  // instead of paying an L2 round trip each under the phase-2 TMA traffic.
  LossStaged stg;
  stg.col = nullptr;
  stg.mk = nullptr;
  if (stage) {
    const int nterms = S.n_prob[1], nslots = S.n_prob[2];
    const int col_bytes = nslots * n * 4;
    if ((n & 15) == 0 && col_bytes + nterms * n <= stage_bytes) {
      float4* dcol = reinterpret_cast<float4*>(stage);
      const float4* scol = reinterpret_cast<const float4*>(L.pcol + (long long)bh * kMaxSlots * n);
      for (int i = tid; i < (col_bytes >> 4); i += 128) dcol[i] = __ldcg(scol + i);
      uint4* dmk = reinterpret_cast<uint4*>(stage + col_bytes);
      const int mvec = n >> 4;
      for (int i = tid; i < nterms * mvec; i += 128) {
        const int t = i / mvec, v = i - t * mvec;
        dmk[i] = __ldg(reinterpret_cast<const uint4*>(L.masks + (long long)S.sterms[t].mask * n) + v);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      stg.col = reinterpret_cast<const float*>(stage);
      stg.mk = stage + col_bytes;
    }
  }
  if (per <= 8)
    loss_problem_loop<8, true>(L, S.sterms, S.stok, S.pcode, S.prob_loss, dpx, n_prob, part * 4 + warp, 4 * nparts,
                               lane, bh, h, heads, n, per, stg, pre, have_pre && !stg.col);
  else {
    LossRaw<40> none;
    loss_problem_loop<40, false>(L, S.sterms, S.stok, S.pcode, S.prob_loss, dpx, n_prob, part * 4 + warp, 4 * nparts,
                                 lane, bh, h, heads, n, per, stg, none, false);
  }
  LOSS_STAMP(1);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (tid == 0) {                                          // fixed summation order => deterministic loss
    float acc = 0.f;
    for (int i = 0; i < n_prob; ++i)
      if (((i >> 2) % nparts) == part) acc += S.prob_loss[i];
    if (nparts == 1) {
      L.loss_part[bh] = acc;
      if (ready) ready[bh] = 0;                            // hand-shake counters end the launch at zero
    } else if (nparts == 2) {
      // two parts: loss_part[bh] was zeroed by part 0 before the hand-shake; 0 + a + b is the same float in either
      // arrival order, so two fire-and-forget atomics replace the store / fence / ticket / re-read chain
      atomicAdd(L.loss_part + bh, acc);
      if (defer_ticket) {
        // the caller consumes the ticket later (loss_ticket_reset), so that this thread's warp - and everybody behind
        // the next barrier - does not sit out the atomic's round trip here
        ticket = atomicAdd(done + bh, 1);
      } else if (atomicAdd(done + bh, 1) == 1) {
        done[bh] = 0;                                      // both parts are past their wait on `ready` by now
        if (ready) ready[bh] = 0;
      }
    } else {
      __stcg(partials + bh * nparts + part, acc);
      __threadfence();
      if (atomicAdd(done + bh, 1) == nparts - 1) {
        __threadfence();
        float tot = 0.f;
        for (int q = 0; q < nparts; ++q) tot += __ldcg(partials + bh * nparts + q);
        L.loss_part[bh] = tot;
        done[bh] = 0;                                      // every part is past its wait on `ready` by now
        if (ready) ready[bh] = 0;
      }
    }
  }
  LOSS_STAMP(2);
  return ticket;
}

// everything at once (unfused kernel: the last CTA of the (image, head) does it all)
__device__ __forceinline__ void xattn_loss_reduce(const XattnLoss& L, float* scratch, float* s_red, int tid, int b,
                                                  int h, int heads, int bh, int n) {
  loss_zero(L, bh, n, 0, n, tid);
  __threadfence_block();
  loss_stage(L, scratch, tid, b);       // its barriers also order the zero-fill before the atomics
  LossRaw<8> none;
  loss_run(L, scratch, tid, h, heads, bh, n, 0, 1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, false, none, false);
  (void)s_red;
}

template <int DPB, int D16>
__global__ void __launch_bounds__(192, 1)
xattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ XattnParams p) {
  using Cfg = XattnCfg<DPB, D16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + Cfg::Q_BYTES;
  uint8_t* sP = sV + 2 * Cfg::V_ATOM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* ld_full = bars;
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* o_full = bars + 3;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr + 1);
  float* s_red = reinterpret_cast<float*>(s_flag + 1);  // 4 floats

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / p.heads, h = bh % p.heads;

  if (warp == 1) {
    if (elect_one()) {
      mbar_init(ld_full, 1);
      mbar_init(s_full, 1);
      mbar_init(p_full, 4);
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
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(ld_full, 2 * Cfg::Q_BYTES + 2 * Cfg::V_ATOM);
      for (int a = 0; a < DPB; ++a) {
        tma_load_3d(sQ + a * 16384, &tmQ, ld_full, a * 64, qt * 128, bh);
        tma_load_3d(sK + a * 16384, &tmK, ld_full, a * 64, 0, bh);
      }
      for (int a = 0; a < 2; ++a) tma_load_3d(sV + a * Cfg::V_ATOM, &tmVt, ld_full, a * 64, 0, bh);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_f16(128, 128);
    constexpr uint32_t idesc_o = make_idesc_f16(128, D16);
    mbar_wait(ld_full, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int a = 0; a < DPB; ++a) {
        const uint64_t ad = make_desc_k_sw128(smem_u32(sQ) + a * 16384);
        const uint64_t bd = make_desc_k_sw128(smem_u32(sK) + a * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tS, ad + k * 2, bd + k * 2, idesc_s, (a | k) ? 1u : 0u);
      }
      tc_commit(s_full);
    }
    __syncwarp();
    mbar_wait(p_full, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const uint64_t ad = make_desc_k_sw128(smem_u32(sP) + a * 16384);
        const uint64_t bd = make_desc_k_sw128(smem_u32(sV) + a * Cfg::V_ATOM);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tO, ad + k * 2, bd + k * 2, idesc_o, (a | k) ? 1u : 0u);
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane_id();
    const int tid = threadIdx.x - 64;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int qrow = qt * 128 + r;
    const bool ok = qrow < p.nq;
    mbar_wait(s_full, 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      if (c0 >= p.nk) break;
      uint32_t v[32];
      tmem_ld_x32(tS + lane_off + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < p.nk) m = fmaxf(m, __uint_as_float(v[i]));
    }
    m *= p.scale_log2;
    float l = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      if (c0 >= p.nk) break;
      uint32_t v[32];
      tmem_ld_x32(tS + lane_off + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < p.nk) l += exp2f(__uint_as_float(v[i]) * p.scale_log2 - m);
    }
    const float inv_l = 1.f / l;
    __half* prow = (p.probs && ok) ? p.probs + ((long long)bh * p.nq + qrow) * p.nk : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t pk[16];
      if (c0 < p.nk) {
        uint32_t v[32];
        tmem_ld_x32(tS + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = (c0 + 2 * i < p.nk) ? exp2f(__uint_as_float(v[2 * i]) * p.scale_log2 - m) * inv_l : 0.f;
          const float bb = (c0 + 2 * i + 1 < p.nk) ? exp2f(__uint_as_float(v[2 * i + 1]) * p.scale_log2 - m) * inv_l : 0.f;
          pk[i] = pack_h2(a, bb);
        }
        if (prow) {
          const __half* hp = reinterpret_cast<const __half*>(pk);
          for (int i = 0; i < 32 && c0 + i < p.nk; ++i) prow[c0 + i] = hp[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
      uint8_t* atom = sP + (c0 >> 6) * 16384;
      const int ch0 = (c0 & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(atom + sw128_offset(r, ch0 + q)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
    // columns of P needed outside: read back this thread's own row from shared memory
    auto p_at = [&](int tok) -> float {
      const uint8_t* atom = sP + (tok >> 6) * 16384;
      const __half* chunk = reinterpret_cast<const __half*>(atom + sw128_offset(r, (tok & 63) >> 3));
      return __half2float(chunk[tok & 7]);
    };
    if (ok && p.save_tok) {
      const int tok = p.save_tok[b];
      if (tok >= 0) p.probs_tok[(long long)bh * p.nq + qrow] = __float2half_rn(p_at(tok));
    }
    if (ok && p.has_loss) {
      for (int s = 0; s < kMaxSlots; ++s) {
        const int tok = p.L.slot_tok[b * kMaxSlots + s];
        if (tok < 0) break;
        p.L.pcol[((long long)bh * kMaxSlots + s) * p.nq + qrow] = p_at(tok);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(p_full);

    // ---------------- O epilogue
    mbar_wait(o_full, 0);
    tc_fence_after();
    __half* orow = p.out + ((long long)b * p.nq + qrow) * p.ldo + h * p.d;
#pragma unroll 1
    for (int c0 = 0; c0 < D16; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tO + lane_off + c0, v);
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
    if (ok && p.lse2) p.lse2[(long long)bh * p.nq_alloc + qrow] = m + log2f(l);
    tc_fence_before();

    // ---------------- guidance loss: last CTA of this (image, head) reduces
    if (p.has_loss) {
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        const int tiles = gridDim.x;
        const int old = atomicAdd(&p.L.counters[bh], 1);
        *s_flag = (old == tiles - 1);
        if (old == tiles - 1) p.L.counters[bh] = 0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*s_flag) {
        __threadfence();
        xattn_loss_reduce(p.L, reinterpret_cast<float*>(sP), s_red, tid, b, h, p.heads, bh, p.nq);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
