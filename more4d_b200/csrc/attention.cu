// Non-causal softmax(Q K^T / sqrt(d)) V for head_dim 128 on tcgen05 tensor cores with TMEM
// accumulators — the replacement for the reference's `attention()` dispatch
// (MoRe4D/models/wan_transformer4d.py:175-236; flash-attn varlen call at :138-169) in the
// layout the reference uses: q/k/v/out are [B, L, heads, 128] bf16 ("NHD").
//
// One CTA = 256 query rows (two 128-row tiles) of one (batch, head); 384 threads:
//   warps 0-3   softmax warpgroup for query tile 0   (thread r owns row r = TMEM lane r)
//   warps 4-7   softmax warpgroup for query tile 1
//   warp  8     TMA producer: Q once, then K_j / V_j tiles (128 keys x 128) through 2-stage rings
//   warp  9     MMA issuer:  S_t = Q_t K_j^T (SS), O_t += P_t V_j (TS: P read from TMEM)
//   warp 10     TMEM allocator (512 columns: S0 | S1 | O0 | O1; P_t aliases S_t as packed bf16)
// The tensor pipe alternates between the two query tiles (QK0 QK1 | PV0 QK0' PV1 QK1' | ...) so
// the softmax of one tile overlaps the MMAs of the other.
//
// Online softmax with lazy rescaling: the running reference max is only moved (and O / l
// rescaled in TMEM) when the row max grows by more than 2^8 in the exp2 domain; otherwise P is
// computed against the stale max, which is exact after the final 1/l normalisation.
// tcgen05 ops issued by one thread complete in order, so "S_t(j) is ready" implies PV_t(j-1)
// has finished: the rescale needs no extra barrier.
//
// Keys beyond k_lens[b] are masked (flash-attn varlen semantics, t4d:122-124); rows beyond Lq
// and keys beyond Lk are zero-filled by TMA and never stored / masked.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace m4d {

constexpr int A_BQ = 128;        // query rows per tile
constexpr int A_NQ = 2;          // query tiles per CTA
constexpr int A_BKV = 128;       // keys per tile
constexpr int A_D = 128;
constexpr int A_KS = 2, A_VS = 2;
constexpr int A_TILE_BYTES = 128 * 128 * 2;   // 32 KB, stored as two [128 x 64] SW128 halves
constexpr int A_HALF_BYTES = 128 * 64 * 2;    // 16 KB
constexpr int A_THREADS = 384;         // 8 softmax warps + TMA + MMA + TMEM allocator
constexpr int A_THREADS_SPLIT = 608;   // MODE 2: 16 softmax warps (two per 32-row group) + the same three
constexpr int A_XCHG_BYTES = 2 * 4 * 2 * 32 * 4;   // MODE 2: one float per (tile, row group, half, lane)
constexpr int A_SMEM_BYTES = (A_NQ + A_KS + A_VS) * A_TILE_BYTES + 256 + 1024 + A_XCHG_BYTES;

constexpr float A_RESCALE_THRESHOLD = 8.0f;   // log2 units
constexpr int A_PERSIST_MAX_KEYS = 2048;      // key ranges up to 16 tiles run on persistent CTAs
constexpr int A_DEFAULT_PP = 2;               // pairs (of 8) whose 2^x runs on the FMA pipe
constexpr int A_DEFAULT_MODE = 1;             // 1: sum-guarded speculative reference
constexpr float A_SUM_GUARD = 65536.0f;       // MODE 1: a half tile whose row sum reaches 2^16 moves the reference

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// (d0, d1) = (a0, a1) * b + c     — one FFMA2
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1,
                                     float c0, float c1) {
  asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1) {
  asm("{ .reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}
// 2^(s*c + neg_mc) for a pair on the FMA/ALU pipes instead of the MUFU.  The argument is clamped to
// [-126, 128] inside the multiply-add that forms it: u = sat((x + 126) / 254) in [0, 1] comes out of
// fma.rn.sat with pre-divided constants (cn = c / 254, bn = (neg_mc + 126) / 254), a round-down
// fma against the magic number 1.5 * 2^23 splits x = u * 254 - 126 into floor and fraction, a
// degree-3 minimax polynomial gives 2^frac (rel. error ~1e-4, far below the bf16 rounding of P) and
// the exponent is re-inserted with one integer multiply-add.  Saturation matters: against a stale
// reference x can be anything; x >= 128 must come out as +inf / huge (so that the row-sum guard
// fires) and x <= -126 as ~0 — the unclamped integer exponent insert wraps around instead.
// 11 instructions per pair; offloads the 16/clk/SM MUFU, which is otherwise co-critical with the
// tensor pipe (128 keys x 256 rows = 2048 MUFU clocks per step and SM sub-partition = the MMA time).
__device__ __forceinline__ void ex2_poly2(float s0, float s1, float cn, float bn, float& r0, float& r1) {
  const float kM = 12582912.0f - 126.0f;         // 1.5 * 2^23 - 126
  float u0, u1, t0, t1;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u0) : "f"(s0), "f"(cn), "f"(bn));
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u1) : "f"(s1), "f"(cn), "f"(bn));
  asm("fma.rm.ftz.f32 %0, %1, %2, %3;" : "=f"(t0) : "f"(u0), "f"(254.0f), "f"(kM));   // floor(x) + 1.5 * 2^23
  asm("fma.rm.ftz.f32 %0, %1, %2, %3;" : "=f"(t1) : "f"(u1), "f"(254.0f), "f"(kM));
  float g0, g1, f0, f1, p0, p1;
  fma2(g0, g1, t0, t1, -1.0f, -1.0f, kM, kM);                      // -126 - floor(x), exact
  fma2(f0, f1, u0, u1, 254.0f, 254.0f, g0, g1);                    // frac = x - floor(x) in [0, 1)
  fma2(p0, p1, f0, f1, 0.077119089663028717f, 0.077119089663028717f, 0.227564394474029541f,
       0.227564394474029541f);
  fma2(p0, p1, p0, p1, f0, f1, 0.695146143436431885f, 0.695146143436431885f);
  fma2(p0, p1, p0, p1, f0, f1, 1.0f, 1.0f);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// Which logit pairs take the polynomial 2^x: PP of every 8, spread evenly so that the MUFU queue and
// the FMA pipe stay busy side by side (PP >= 8: the contiguous tail pattern of round 1, PP - 8 of 8).
template <int PP>
__device__ __forceinline__ constexpr bool poly_pair(int q) {
  return PP == 0 ? false
       : PP == 1 ? (q & 7) == 7
       : PP == 2 ? (q & 3) == 3
       : PP == 3 ? ((q & 7) == 2 || (q & 7) == 5 || (q & 7) == 7)
       : PP == 4 ? (q & 1) == 1
       : (q & 7) >= 16 - PP;   // PP = 10: pairs 6, 7 of every 8
}

// p = 2^(s*c + neg_mc) for the logit pairs [Q0, Q0 + NQ) of s (KEY0/2 + q = pair index inside the
// 128-key tile, which fixes the MUFU / polynomial assignment per key column), packed bf16x2 into
// pk[0 .. NQ); row-sum partials accumulate pairwise (FADD2) in pair order.  The instruction order
// is ptxas's: it interleaves ~4 FMA-pipe instructions per MUFU.EX2 and keeps the consumers of a
// pair one pair behind its EX2s whatever order the source uses (measured on the SASS).
template <int PP, int Q0, int NQ, int KEY0 = 0>
__device__ __forceinline__ void exp_pipe(const uint32_t* s, uint32_t* pk, float c, float neg_mc,
                                         float& l0, float& l1) {
  const float cn = c * (1.0f / 254.0f), bn = (neg_mc + 126.0f) * (1.0f / 254.0f);
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int q = Q0 + i;
    const float s0 = __uint_as_float(s[2 * q]), s1 = __uint_as_float(s[2 * q + 1]);
    float e0, e1;
    if (PP != 0 && poly_pair<PP>(KEY0 / 2 + q)) {
      ex2_poly2(s0, s1, cn, bn, e0, e1);
    } else {
      float x0, x1;
      fma2(x0, x1, s0, s1, c, c, neg_mc, neg_mc);
      e0 = fast_exp2(x0);
      e1 = fast_exp2(x1);
    }
    add2(l0, l1, e0, e1);
    pk[i] = pack_bf16(e0, e1);
  }
}

// max over 32 logits (FMNMX3: two per instruction), keys >= valid excluded
__device__ __forceinline__ float max32(const uint32_t* s, int valid, float m) {
  float a = m, b = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    a = max3(a, i < valid ? __uint_as_float(s[i]) : -INFINITY, i + 1 < valid ? __uint_as_float(s[i + 1]) : -INFINITY);
    b = max3(b, i + 2 < valid ? __uint_as_float(s[i + 2]) : -INFINITY,
             i + 3 < valid ? __uint_as_float(s[i + 3]) : -INFINITY);
  }
  return fmaxf(a, b);
}

struct AttnParams {
  bf16* out;
  long long out_stride_b, out_stride_l;   // elements
  int Lq, Lk, heads;
  const int* k_lens;                      // [B] or null
  float scale_log2;                       // softmax_scale * log2(e)
  int accumulate;                         // out = bf16(out + bf16(o))  (summed cross-attention)
  // Two independently normalised attentions in ONE launch (WanI2VCrossAttention, t4d:533-552: text
  // and CLIP-image context, outputs summed in bf16): keys [0, seg_tiles*128) and [seg_tiles*128, Lk)
  // are separate softmax segments.  At the segment boundary the softmax warps normalise and store
  // the first segment's O, start over (reference max, row sum, accumulator), and the final
  // epilogue adds the second segment onto the stored bf16 — one Q load, one prologue, one epilogue
  // pair instead of two launches that each run 3-4 key tiles.  0 = one segment.
  int seg_tiles;
  // PERSIST kernels: work items (q block, head, batch), q block fastest; n_items = n_qblk * heads * B
  int n_qblk, n_items;
  // scatter epilogue (sequence-parallel exchange fused into the attention epilogue): query row l
  // is stored to out_scatter[l / scatter_rows] at row l % scatter_rows — the destinations are
  // the ranks' receive buffers (peer memory), one launch serves all of them
  bf16* out_scatter[8];
  int scatter_rows;                       // 0 = plain output
};

// PP: pairs of every 8 whose exponential runs on the FMA pipe (ex2_poly2); MODE 0: exact row max
// before the exponentials, MODE 1: sum-guarded speculative reference (see the softmax section).  P is handed to the tensor pipe in two 64-key halves.
// Measured variants that did NOT help on the B200 and were removed again (round 1,
// profiles/attn_variant_sweep_r01.log): one mbarrier arrival per warp instead of per thread
// (-4 %), P in 1 or 4 slices, 64-key steps with a double-buffered S (-25 %: N=64 QK MMAs are
// shared-memory bound), two threads per row (16 softmax warps share the MUFU: -9 %).
// PERSIST: one CTA per SM loops over work items (q block, head, batch).  For the short key ranges of
// the cross-attention (7 key tiles: ~10 us of tensor work) a CTA per item spent ~26 us on launch,
// barrier / TMEM set-up, the first loads and the epilogue with the tensor pipe idle (36 us per CTA,
// 425 TF/s).  Here the next item's Q / K / V loads and its first QK^T are issued while the softmax
// warps are still in the previous item's epilogue; barrier phases run on (item, tile) counters.
template <int PP, int MODE, bool PERSIST = false>
__global__ void __launch_bounds__(MODE == 2 ? A_THREADS_SPLIT : A_THREADS, 1)
attn_fwd_d128_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + A_NQ * A_TILE_BYTES;
  uint8_t* sV = sK + A_KS * A_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + A_VS * A_TILE_BYTES);
  uint64_t* q_full = bars;              // 1
  uint64_t* k_full = q_full + 1;        // A_KS
  uint64_t* k_empty = k_full + A_KS;    // A_KS
  uint64_t* v_full = k_empty + A_KS;    // A_VS
  uint64_t* v_empty = v_full + A_VS;    // A_VS
  uint64_t* s_full = v_empty + A_VS;    // 2
  uint64_t* p_ready = s_full + 2;       // per tile t: [4t + 0/1] P halves ready, [4t + 2] PV of the first half done
  uint64_t* o_final = p_ready + 8;      // 2
  uint64_t* q_empty = o_final + 2;      // 1   PERSIST: every QK^T of the item has been executed
  uint64_t* o_free = q_empty + 1;       // 2   PERSIST: the item's epilogue has drained O_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // MODE 2 partner exchange

  constexpr int NSW = (MODE == 2) ? 16 : 8;       // softmax warps; then TMA, MMA, TMEM-allocator warps
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(!(PERSIST && MODE == 2), "the split-column variant is not persistent");
  const int n_iter = PERSIST ? (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                   static_cast<int>(gridDim.x)
                             : 1;
  auto decode = [&](int it, int& q_blk_, int& head_, int& b_) {
    if (PERSIST) {
      const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      q_blk_ = item % p.n_qblk;
      const int r = item / p.n_qblk;
      head_ = r % p.heads;
      b_ = r / p.heads;
    } else {
      q_blk_ = blockIdx.x;
      head_ = blockIdx.y;
      b_ = blockIdx.z;
    }
  };
  int q_blk, head, b;
  decode(0, q_blk, head, b);

  int kv_len = p.Lk;
  if (!PERSIST && p.k_lens != nullptr) {           // persistent launches take no k_lens (checked on the host)
    int kl = p.k_lens[b];
    kv_len = kl < kv_len ? kl : kv_len;
  }
  if (kv_len < 1) kv_len = 1;
  const int n_kv = (kv_len + A_BKV - 1) / A_BKV;

  if (warp == NSW && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == NSW + 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < A_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < A_VS; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      for (int i = 0; i < 4; ++i) mbar_init(&p_ready[4 * t + i], i == 2 ? 1 : 128);   // [2]: pv_half, by commit
      mbar_init(&o_final[t], 1);
      mbar_init(&o_free[t], 128);
    }
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == NSW + 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= NSW) {
    // ------------------------------------------------------------------ data movement + MMA
    if (MODE != 2) reg_dec<80>();
    if (warp == NSW && lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        if (PERSIST) {
          decode(it, q_blk, head, b);
          mbar_wait(q_empty, (it & 1) ^ 1);        // the previous item's QK^T MMAs have read Q
        }
        mbar_arrive_expect_tx(q_full, A_NQ * A_TILE_BYTES);
        for (int t = 0; t < A_NQ; ++t)
          for (int h = 0; h < 2; ++h)
            tma_load_4d(sQ + t * A_TILE_BYTES + h * A_HALF_BYTES, &tmQ, q_full, h * 64, head,
                        q_blk * (A_NQ * A_BQ) + t * A_BQ, b);
        for (int j = 0; j < n_kv; ++j) {
          const int g = it * n_kv + j;             // running tile counter: ring slot and parity
          const int sk = g % A_KS, sv = g % A_VS;
          mbar_wait(&k_empty[sk], ((g / A_KS) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[sk], A_TILE_BYTES);
          for (int h = 0; h < 2; ++h)
            tma_load_4d(sK + sk * A_TILE_BYTES + h * A_HALF_BYTES, &tmK, &k_full[sk], h * 64, head,
                        j * A_BKV, b);
          mbar_wait(&v_empty[sv], ((g / A_VS) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[sv], A_TILE_BYTES);
          for (int h = 0; h < 2; ++h)
            tma_load_4d(sV + sv * A_TILE_BYTES + h * A_HALF_BYTES, &tmV, &v_full[sv], h * 64, head,
                        j * A_BKV, b);
        }
      }
    } else if (warp == NSW + 1) {
      // The whole warp runs this loop converged and ONE elected lane issues: operands then live
      // in uniform registers and every descriptor is {lo + compile-time offset, constant hi}.
      // With `if (lane == 0)` the compiler wraps each tcgen05.mma in an R2UR/ELECT loop and
      // rebuilds the descriptor (~14 SASS instructions per 64-clock MMA): the issue rate was
      // co-critical with the tensor pipe.
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);   // V is MN-major (d contiguous)
      // K-major SW128 tiles (Q, K): LBO 16 B, SBO 1024 B.  V (MN-major): 64-wide d chunks are
      // 16 KB apart (LBO), 8-key groups 1 KB apart (SBO).
      constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t q_lo = ((smem_u32(sQ) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t k_lo = ((smem_u32(sK) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t v_lo = ((smem_u32(sV) >> 4) & 0x3FFF) | ((A_HALF_BYTES >> 4) << 16);
      const uint32_t tS0 = tmem_base, tO0 = tmem_base + 256;
      auto dp = [](uint32_t lo, uint32_t hi) { return (static_cast<uint64_t>(hi) << 32) | lo; };

      auto issue_qk = [&](int t, int sk) {               // elected lane only
        const uint32_t a_lo = q_lo + t * (A_TILE_BYTES >> 4), b_lo = k_lo + sk * (A_TILE_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < A_D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * (A_HALF_BYTES >> 4) + (kk & 3) * 2;
          umma_ss(tS0 + t * 128, dp(a_lo + off, HI), dp(b_lo + off, HI), idesc_qk, kk != 0);
        }
      };
      // O_t += P_t[:, slice] V[slice, :] — P arrives in NS key-slices so the first PV MMAs
      // overlap the exponentials of the later slices
      constexpr int NS = 2;
      auto issue_pv_tile = [&](int t, int sv, int j, int g, int it, bool leader) {   // whole warp
        const uint32_t b_lo = v_lo + sv * (A_TILE_BYTES >> 4);
        // PERSIST: the first PV of an item overwrites O_t — the previous item's epilogue must be done
        if (PERSIST && j == 0 && it > 0) mbar_wait(&o_free[t], (it - 1) & 1);
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
          mbar_wait(&p_ready[4 * t + sl], g & 1);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int k4 = 0; k4 < 8 / NS; ++k4) {
              const int kk = sl * (8 / NS) + k4;
              umma_ts(tO0 + t * 128, tS0 + t * 128 + kk * 8, dp(b_lo + kk * (2048 >> 4), HI), idesc_pv,
                      !((j == 0 || j == p.seg_tiles) && kk == 0));
            }
            // MODE 1: a softmax warp whose SECOND P half needs a new reference must rescale O
            // after these MMAs and before the next ones (rare; nobody waits otherwise)
            if (MODE == 1 && sl == 0) umma_commit(&p_ready[4 * t + 2]);
          }
          __syncwarp();
        }
      };

      const bool leader = elect_one();
      for (int it = 0; it < n_iter; ++it) {
        const int g0 = it * n_kv;
        mbar_wait(q_full, it & 1);
        mbar_wait(&k_full[g0 % A_KS], (g0 / A_KS) & 1);
        tc_fence_after();
        if (leader) {
          // S_t is free: the previous item's last PV_t was issued before this and executes in order
          issue_qk(0, g0 % A_KS);
          umma_commit(&s_full[0]);
          issue_qk(1, g0 % A_KS);
          umma_commit(&s_full[1]);
          umma_commit(&k_empty[g0 % A_KS]);
        }
        __syncwarp();
        for (int j = 0; j < n_kv; ++j) {
          const int g = g0 + j;
          const int sv = g % A_VS;
          const bool last = (j + 1 == n_kv);
          const int sk = (g + 1) % A_KS;
          if (PERSIST && last && leader) umma_commit(q_empty);   // all QK^T of this item are issued
          mbar_wait(&v_full[sv], (g / A_VS) & 1);
          issue_pv_tile(0, sv, j, g, it, leader);
          if (!last) mbar_wait(&k_full[sk], ((g + 1) / A_KS) & 1);
          tc_fence_after();
          if (leader) {
            if (last) {
              umma_commit(&o_final[0]);
            } else {
              issue_qk(0, sk);
              umma_commit(&s_full[0]);
            }
          }
          __syncwarp();
          issue_pv_tile(1, sv, j, g, it, leader);
          if (leader) {
            umma_commit(&v_empty[sv]);
            if (last) {
              umma_commit(&o_final[1]);
            } else {
              issue_qk(1, sk);
              umma_commit(&s_full[1]);
              umma_commit(&k_empty[sk]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (MODE == 2) {
    // ------------------------------------------------------------------ softmax, split columns
    // MODE 2.  With one warp per 32 rows the softmax of a tile is a single in-order instruction
    // stream: ~560 instructions (MODE 1, PP 2) took ~1700 clocks against 768 clocks of MUFU work
    // and 1024 clocks of MMA time per tile (profiles/attn_r02a.md: the softmax warps are busy 62 %
    // of the step, the tensor pipe idles while both tiles are in their softmax).  Here TWO warps
    // share a row group: warp (t, half 0, quad) owns key columns 0..63 of its 32 rows, warp
    // (t, half 1, quad) columns 64..127 — two independent instruction streams per sub-partition
    // and tile hide each other's latencies.  The sum-guarded reference of MODE 1 makes that cheap:
    // no row max, hence no per-tile exchange between the partners; each half's row sum is the
    // guard of that half and ONE named barrier with an OR-reduction (barrier.red.or) per tile tells
    // both warps whether any of their 64 threads tripped it.  If so (always at j = 0, rarely
    // later) both take the exact path together: publish the half-row maxima of the rows that
    // tripped, move those rows' reference to the same value in both partners, rescale O (each
    // warp its 64 output columns) / l, redo both halves from S (intact: the barrier precedes the
    // P stores).  l is kept as two partial sums and combined once in the epilogue.
    const int t = warp >> 3;                       // query tile
    const int half = (warp >> 2) & 1;              // key-column half of the tile
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // row in tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * 128;
    const uint32_t tSh = tS + half * 64;           // my 64 logit columns
    const uint32_t tPh = tS + half * 32;           // my 32 packed-P columns
    const uint32_t tOh = tmem_base + lane_base + 256 + t * 128 + half * 64;   // my 64 output columns
    const float c = p.scale_log2;
    const int bar_id = 1 + t * 4 + quad;           // named barrier of the two partner warps
    float* my_x = xchg + ((t * 4 + quad) * 2 + half) * 32 + lane;
    float* partner_x = xchg + ((t * 4 + quad) * 2 + (half ^ 1)) * 32 + lane;
    uint32_t pr;       // opaque to the compiler, otherwise it re-derives the address late (and the arrive with it)
    asm volatile("mov.u32 %0, %1;" : "=r"(pr) : "r"(smem_u32(&p_ready[4 * t + half])));
    float m_used = -INFINITY;                      // reference, in logit units; identical in both partners
    float l_part = 0.f;                            // row sum over my key columns

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[64];
      uint32_t pk[32];
      const int valid = kv_len - j * A_BKV - half * 64;      // keys of my half that exist (may be <= 0)
      float l0 = 0.f, l1 = 0.f;
      tmem_ld32(tSh, s);
      tmem_ld_wait();                              // first 32 logits have landed
      tmem_ld32(tSh + 32, s + 32);                 // in flight while they are processed
      reg_fence32(s);
      if (valid < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) s[i] = 0xFF800000u;      // -inf
      }
      float neg_mc = -m_used * c;
      if (half == 0) exp_pipe<PP, 0, 16, 0>(s, pk, c, neg_mc, l0, l1);
      else           exp_pipe<PP, 0, 16, 64>(s, pk, c, neg_mc, l0, l1);
      tmem_ld_wait();
      reg_fence32(s + 32);
      if (valid < 64) {
#pragma unroll
        for (int i = 32; i < 64; ++i)
          if (i >= valid) s[i] = 0xFF800000u;
      }
      if (half == 0) exp_pipe<PP, 16, 16, 0>(s, pk + 16, c, neg_mc, l0, l1);
      else           exp_pipe<PP, 16, 16, 64>(s, pk + 16, c, neg_mc, l0, l1);
      const bool tripped = (j == 0) || !(l0 + l1 < A_SUM_GUARD);   // j = 0: no reference yet
      uint32_t any;
      asm volatile("{\n\t.reg .pred pi, po;\n\t"
                   "setp.ne.u32 pi, %1, 0;\n\t"
                   "barrier.cta.red.or.pred po, %2, 64, pi;\n\t"
                   "selp.u32 %0, 1, 0, po;\n\t}\n"
                   : "=r"(any) : "r"(static_cast<uint32_t>(tripped)), "r"(bar_id) : "memory");
      if (any) {
        // exact path, both partner warps.  Nothing of this tile has been stored, so S is intact;
        // "S_t(j) ready" implies PV_t(j-1) has finished, so O may be rescaled.
        float mxh = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t o[32];
          tmem_ld32(tSh + cc * 32, o);
          tmem_ld_wait();
          mxh = max32(o, valid - cc * 32, mxh);
        }
        *my_x = tripped ? mxh : -INFINITY;         // only rows that tripped THEMSELVES move (per-row decision)
        named_bar_sync(bar_id, 64);
        const float m_new = fmaxf(m_used, fmaxf(tripped ? mxh : -INFINITY, *partner_x));
        const float f = (j == 0) ? 0.f : fast_exp2((m_used - m_new) * c);   // 1 for rows that did not move
        m_used = m_new;
        l_part *= f;
        if (j > 0) {
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {         // my 64 output columns; every lane takes part
            uint32_t o[32];
            tmem_ld32(tOh + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(tOh + cc * 32, o);
          }
          tmem_st_wait();
        }
        neg_mc = -m_used * c;
        tmem_ld32(tSh, s);
        tmem_ld32(tSh + 32, s + 32);
        tmem_ld_wait();
        reg_fence32(s);
        reg_fence32(s + 32);
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) s[i] = 0xFF800000u;
        }
        l0 = 0.f;
        l1 = 0.f;
        if (half == 0) exp_pipe<PP, 0, 32, 0>(s, pk, c, neg_mc, l0, l1);
        else           exp_pipe<PP, 0, 32, 64>(s, pk, c, neg_mc, l0, l1);
        named_bar_sync(bar_id, 64);                // the partner has re-read its S columns: P may overwrite them
      }
      tmem_st32(tPh, pk);
      tmem_st_wait();
      tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pr) : "memory");
      l_part += l0 + l1;
    }

    // ---- epilogue: O / l -> bf16 -> global, 64 output columns per warp
    *my_x = l_part;
    named_bar_sync(bar_id, 64);
    const float inv_l = 1.0f / (half == 0 ? l_part + *partner_x : *partner_x + l_part);   // same order in both partners
    mbar_wait(&o_final[t], 0);
    tc_fence_after();
    const int q_row = q_blk * (A_NQ * A_BQ) + t * A_BQ + row;
    const bool row_ok = q_row < p.Lq;
    bf16* obase = p.out;
    int o_row = q_row;
    if (p.scatter_rows > 0 && row_ok) {
      const int dst = q_row / p.scatter_rows;
      obase = p.out_scatter[dst];
      o_row = q_row - dst * p.scatter_rows;
    }
    bf16* orow = obase + static_cast<long long>(b) * p.out_stride_b +
                 static_cast<long long>(o_row) * p.out_stride_l + head * A_D + half * 64;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t o[32];
      tmem_ld32(tOh + cc * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[q * 8 + e]) * inv_l;
          uint4* dst = reinterpret_cast<uint4*>(orow + cc * 32 + q * 8);
          if (p.accumulate) {
            const uint4 prev = *dst;
            const uint32_t w[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[2 * e] = bf16_round(v[2 * e]) + __uint_as_float(w[e] << 16);
              v[2 * e + 1] = bf16_round(v[2 * e + 1]) + __uint_as_float(w[e] & 0xFFFF0000u);
            }
          }
          uint4 ov;
          ov.x = pack_bf16(v[0], v[1]);
          ov.y = pack_bf16(v[2], v[3]);
          ov.z = pack_bf16(v[4], v[5]);
          ov.w = pack_bf16(v[6], v[7]);
          *dst = ov;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    // Per 128-key step and query tile the chain  S ready -> P ready -> PV -> QK(next) -> S ready
    // bounds the step (TMEM is full, so S(j+1) cannot be produced while P(j) is live); the
    // softmax part of it is kept as short as the MUFU allows.
    //
    // MODE 1 — sum-guarded speculative reference.  Softmax is invariant to the reference r in
    // p = 2^(s*c - r) as long as nothing over- or underflows, so the row max need not be known:
    // the exponentials of tile j start against the reference left by the earlier tiles the moment
    // the first 32 logits arrive from TMEM (no max pass: 64 FMNMX3 and ~235 clocks per tile less),
    // and the row sum they produce anyway is the guard: a half tile whose sum stays below 2^16
    // (no inf / NaN either) proves every p < 2^16 and is handed to the tensor pipe as it is.
    // Otherwise (always at j = 0, rarely later) the rows concerned take the exact path: reload
    // S — still intact in TMEM —, reduce the true max, move the reference, rescale O / l, redo the
    // half.  For the SECOND half the first one is already with the tensor pipe at the old
    // reference: the warp waits for those MMAs (`pv_half`, committed by the MMA warp on every step
    // and waited on only here) before it rescales O.  The reference never exceeds the running row
    // max, so nothing underflows that the textbook form would keep.
    // MODE 0 — exact: max pass over the tile first, lazy rescale (threshold 2^8), then the
    // exponentials; kept as the baseline of the measurement.
    // In both modes a row's result depends on its own logits only (per-row decisions, fixed
    // MUFU / polynomial assignment per key column, fixed summation order): any query chunking —
    // the sequence-parallel launches — is bit-identical.
    reg_inc<208>();
    const int t = warp >> 2;                       // query tile
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // row in tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * 128;
    const uint32_t tO = tmem_base + lane_base + 256 + t * 128;
    const float c = p.scale_log2;
    uint32_t pr0;      // opaque to the compiler, otherwise it re-derives the address late (and the arrive with it)
    asm volatile("mov.u32 %0, %1;" : "=r"(pr0) : "r"(smem_u32(&p_ready[4 * t + 0])));
    float m_used = -INFINITY;                      // reference, in logit units (r = m_used * c)
    float l_sum = 0.f;
    bool row_ok = false;                           // per work item (set at the top of the item loop)
    bool row32 = false;                            // the output row is 32-byte aligned
    bf16* orow = nullptr;

    // O *= f, by 32-column chunks (rare path)
    auto rescale_o = [&](float f) {
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t o[32];
        tmem_ld32(tO + cc * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
        tmem_st32(tO + cc * 32, o);
      }
      tmem_st_wait();
    };

    auto locate_rows = [&]() {
      const int q_row = q_blk * (A_NQ * A_BQ) + t * A_BQ + row;
      row_ok = q_row < p.Lq;
      bf16* obase = p.out;
      int o_row = q_row;
      if (p.scatter_rows > 0 && row_ok) {
        const int dst = q_row / p.scatter_rows;
        obase = p.out_scatter[dst];
        o_row = q_row - dst * p.scatter_rows;
      }
      orow = obase + static_cast<long long>(b) * p.out_stride_b +
             static_cast<long long>(o_row) * p.out_stride_l + head * A_D;
      row32 = (reinterpret_cast<uintptr_t>(orow) & 31) == 0;
    };
    // O / l -> bf16 -> global (accum: out = bf16(bf16(o) + out)).  All four 32-column TMEM loads are
    // issued before the first wait, and `prev` — the row's previous output for the accumulate forms,
    // loaded by the caller BEFORE it waits for the last PV — keeps the global-load latency out of the
    // epilogue (both were the top stalls of the cross-attention: profiles/cross_attn_r02.md).
    // 256-bit accesses when the row is 32-byte aligned (whole sectors, ptx.cuh).
    auto load_prev = [&](uint32_t* prev) {
      if (!row_ok) return;
      if (row32) {
#pragma unroll
        for (int q = 0; q < 8; ++q) ldg256(orow + q * 16, prev + q * 8);
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const uint4 w = *reinterpret_cast<const uint4*>(orow + q * 8);
          prev[q * 4 + 0] = w.x; prev[q * 4 + 1] = w.y; prev[q * 4 + 2] = w.z; prev[q * 4 + 3] = w.w;
        }
      }
    };
    auto store_o = [&](float inv_l, const uint32_t* prev) {     // prev == nullptr: plain store
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {             // 64 columns at a time: two TMEM loads in flight
        uint32_t o[64];
        tmem_ld32(tO + hf * 64, o);
        tmem_ld32(tO + hf * 64 + 32, o + 32);
        tmem_ld_wait();
        reg_fence32(o);
        reg_fence32(o + 32);
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {            // 16 channels per step
            uint32_t w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v0 = __uint_as_float(o[q * 16 + 2 * e]) * inv_l, v1 = __uint_as_float(o[q * 16 + 2 * e + 1]) * inv_l;
              if (prev != nullptr) {
                v0 = bf16_round(v0) + bf16_lo(prev[hf * 32 + q * 8 + e]);
                v1 = bf16_round(v1) + bf16_hi(prev[hf * 32 + q * 8 + e]);
              }
              w[e] = pack_bf16(v0, v1);
            }
            bf16* dst = orow + hf * 64 + q * 16;
            if (row32) {
              stg256(dst, w);
            } else {
              *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        }
      }
    };

    for (int it = 0; it < n_iter; ++it) {
    if (PERSIST) decode(it, q_blk, head, b);
    locate_rows();
    m_used = -INFINITY;
    l_sum = 0.f;
    const int g0 = it * n_kv;
    for (int j = 0; j < n_kv; ++j) {
      const int g = g0 + j;                        // running tile counter: barrier parities
      mbar_wait(&s_full[t], g & 1);
      tc_fence_after();
      if (__builtin_expect(j > 0 && j == p.seg_tiles, 0)) {
        // segment boundary: "S_t(j) ready" implies PV_t(j-1) has finished, so O holds the whole
        // first segment; the next PV starts a fresh accumulator (MMA warp) — finish this one
        store_o(1.0f / l_sum, nullptr);             // (two segments never come with accumulate: host check)
        m_used = -INFINITY;
        l_sum = 0.f;
      }
      const bool first = (j == 0) || (j == p.seg_tiles);   // no reference yet in this segment
      uint32_t s[128];
      uint32_t ph[32];
      const int valid = kv_len - j * A_BKV;        // keys of this tile that exist
      float l0 = 0.f, l1 = 0.f;
      if (MODE == 1) {
        float neg_mc = -m_used * c;
        bool exact = false;
        if (__builtin_expect(!first, 1)) {
          // ---- first half, speculative
          tmem_ld32(tS + 0, s + 0);
          tmem_ld_wait();                          // chunk 0 has landed
          tmem_ld32(tS + 32, s + 32);              // in flight while chunk 0 is processed
          tmem_ld32(tS + 64, s + 64);
          tmem_ld32(tS + 96, s + 96);
          reg_fence32(s + 0);
          if (__builtin_expect(valid < 32, 0)) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= valid) s[i] = 0xFF800000u;  // -inf
          }
          exp_pipe<PP, 0, 16>(s, ph, c, neg_mc, l0, l1);
          tmem_ld_wait();
          reg_fence32(s + 32);
          reg_fence32(s + 64);
          reg_fence32(s + 96);
          if (__builtin_expect(valid < A_BKV, 0)) {
#pragma unroll
            for (int i = 32; i < 128; ++i)
              if (i >= valid) s[i] = 0xFF800000u;
          }
          exp_pipe<PP, 16, 16>(s, ph + 16, c, neg_mc, l0, l1);
          exact = __any_sync(0xffffffffu, !(l0 + l1 < A_SUM_GUARD));
        } else {
          // a segment's first tile: the whole tile in one TMEM read, row max from registers (the
          // textbook order), first half's exponentials; O and l start from zero
          tmem_ld32(tS + 0, s + 0);
          tmem_ld32(tS + 32, s + 32);
          tmem_ld32(tS + 64, s + 64);
          tmem_ld32(tS + 96, s + 96);
          tmem_ld_wait();
          reg_fence32(s + 0);
          reg_fence32(s + 32);
          reg_fence32(s + 64);
          reg_fence32(s + 96);
          if (valid < A_BKV) {
#pragma unroll
            for (int i = 0; i < 128; ++i)
              if (i >= valid) s[i] = 0xFF800000u;
          }
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            mx[0] = max3(mx[0], __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
            mx[1] = max3(mx[1], __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
            mx[2] = max3(mx[2], __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
            mx[3] = max3(mx[3], __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
          }
          m_used = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          l_sum = 0.f;
          neg_mc = -m_used * c;
          exp_pipe<PP, 0, 32>(s, ph, c, neg_mc, l0, l1);
        }
        if (__builtin_expect(exact, 0)) {
          // exact path for the whole tile: S is intact (nothing of this tile has been stored);
          // "S_t(j) ready" implies PV_t(j-1) has finished, so O may be rescaled
          const bool mine = !(l0 + l1 < A_SUM_GUARD);
          float mxr = -INFINITY;
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t o[32];
            tmem_ld32(tS + cc * 32, o);
            tmem_ld_wait();
            mxr = max32(o, valid - cc * 32, mxr);
          }
          float f = 1.0f;
          if (mine) {
            const float m_new = fmaxf(m_used, mxr);
            f = fast_exp2((m_used - m_new) * c);
            m_used = m_new;
            l_sum *= f;
          }
          rescale_o(f);                            // warp-collective tcgen05 ops: every lane takes part
          neg_mc = -m_used * c;
          tmem_ld32(tS + 0, s + 0);
          tmem_ld32(tS + 32, s + 32);
          tmem_ld_wait();
          reg_fence32(s + 0);
          reg_fence32(s + 32);
          if (valid < 64) {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i >= valid) s[i] = 0xFF800000u;
          }
          l0 = 0.f;
          l1 = 0.f;
          exp_pipe<PP, 0, 32>(s, ph, c, neg_mc, l0, l1);
        }
        tmem_st32(tS, ph);
        tmem_st_wait();
        tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pr0) : "memory");
        // ---- second half
        float h0 = 0.f, h1 = 0.f;
        exp_pipe<PP, 32, 32>(s, ph, c, neg_mc, h0, h1);
        if (__builtin_expect(__any_sync(0xffffffffu, !(h0 + h1 < A_SUM_GUARD)), 0)) {
          // the first half is with the tensor pipe at the old reference: wait for its MMAs,
          // then move the reference (true max of the second half), rescale O / l, redo the half
          const bool mine = !(h0 + h1 < A_SUM_GUARD);
          mbar_wait(&p_ready[4 * t + 2], g & 1);
          tc_fence_after();
          float mxr = -INFINITY;
#pragma unroll 1
          for (int cc = 2; cc < 4; ++cc) {
            uint32_t o[32];
            tmem_ld32(tS + cc * 32, o);
            tmem_ld_wait();
            mxr = max32(o, valid - cc * 32, mxr);
          }
          float f = 1.0f;
          if (mine) {
            const float m_new = fmaxf(m_used, mxr);
            f = fast_exp2((m_used - m_new) * c);
            m_used = m_new;
            l_sum *= f;
            l0 *= f;
            l1 *= f;
          }
          rescale_o(f);
          neg_mc = -m_used * c;
          tmem_ld32(tS + 64, s + 64);
          tmem_ld32(tS + 96, s + 96);
          tmem_ld_wait();
          reg_fence32(s + 64);
          reg_fence32(s + 96);
          if (valid < A_BKV) {
#pragma unroll
            for (int i = 64; i < 128; ++i)
              if (i >= valid) s[i] = 0xFF800000u;
          }
          h0 = 0.f;
          h1 = 0.f;
          exp_pipe<PP, 32, 32>(s, ph, c, neg_mc, h0, h1);
        }
        tmem_st32(tS + 32, ph);
        tmem_st_wait();
        tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+8];" ::"r"(pr0) : "memory");
        l_sum += (l0 + l1) + (h0 + h1);
      } else {
        // ---- MODE 0: exact max first
        tmem_ld32(tS + 0, s + 0);
        tmem_ld32(tS + 32, s + 32);
        tmem_ld32(tS + 64, s + 64);
        tmem_ld32(tS + 96, s + 96);
        tmem_ld_wait();
        reg_fence32(s + 0);
        reg_fence32(s + 32);
        reg_fence32(s + 64);
        reg_fence32(s + 96);
        if (valid < A_BKV) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) s[i] = 0xFF800000u;
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          mx[0] = max3(mx[0], __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mx[1] = max3(mx[1], __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
          mx[2] = max3(mx[2], __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
          mx[3] = max3(mx[3], __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
        }
        const float mxr = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (first) {
          m_used = mxr;
        } else {
          const float m_new = fmaxf(m_used, mxr);
          const bool grow = (m_new - m_used) * c > A_RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, grow)) {
            const float f = grow ? fast_exp2((m_used - m_new) * c) : 1.0f;
            if (grow) m_used = m_new;
            l_sum *= f;
            rescale_o(f);
          }
        }
        const float neg_mc = -m_used * c;
        exp_pipe<PP, 0, 32>(s, ph, c, neg_mc, l0, l1);
        tmem_st32(tS, ph);
        tmem_st_wait();
        tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pr0) : "memory");
        exp_pipe<PP, 32, 32>(s, ph, c, neg_mc, l0, l1);
        tmem_st32(tS + 32, ph);
        tmem_st_wait();
        tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+8];" ::"r"(pr0) : "memory");
        l_sum += l0 + l1;
      }
    }

    // ---- epilogue: O / l -> bf16 -> global
    if (p.accumulate != 0 || p.seg_tiles > 0) {
      uint32_t prev[64];
      load_prev(prev);                             // in flight while the last PV finishes
      mbar_wait(&o_final[t], it & 1);
      tc_fence_after();
      store_o(1.0f / l_sum, prev);
    } else {
      mbar_wait(&o_final[t], it & 1);
      tc_fence_after();
      store_o(1.0f / l_sum, nullptr);
    }
    if (PERSIST) {                                 // O_t may be overwritten by the next item's first PV
      tc_fence_before();
      mbar_arrive(&o_free[t]);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NSW + 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace m4d

using namespace m4d;

static int attention_impl(const void* q, const void* k, const void* v, void* out, int B,
                          int Lq, int Lk, int heads, int head_dim, long long q_stride_b,
                          long long q_stride_l, long long kv_stride_b, long long kv_stride_l,
                          long long out_stride_b, long long out_stride_l, const int* k_lens,
                          float softmax_scale, int accumulate, void* const* out_scatter, int n_scatter,
                          int scatter_rows, int seg_len, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_scatter > 0) {
    M4D_REQUIRE(out_scatter && n_scatter <= 8 && scatter_rows > 0 &&
                    static_cast<long long>(n_scatter) * scatter_rows >= Lq && !accumulate,
                M4D_ERR_BAD_SHAPE);
    for (int i = 0; i < n_scatter; ++i) M4D_REQUIRE(out_scatter[i] && aligned16(out_scatter[i]), M4D_ERR_ALIGN);
    out = out_scatter[0];
  }
  M4D_REQUIRE(q && k && v && out, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && heads > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(head_dim == A_D, M4D_ERR_UNSUPPORTED);          // d = 128 is the Wan2.1 invariant
  M4D_REQUIRE(seg_len >= 0 && seg_len % A_BKV == 0 && seg_len < Lk &&
                  (seg_len == 0 || (k_lens == nullptr && n_scatter == 0 && accumulate == 0)),
              M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(heads <= 65535 && B <= 65535, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(q_stride_l >= heads * A_D && kv_stride_l >= heads * A_D &&
                  out_stride_l >= heads * A_D,
              M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(q_stride_l % 8 == 0 && kv_stride_l % 8 == 0 && out_stride_l % 8 == 0 &&
                  q_stride_b % 8 == 0 && kv_stride_b % 8 == 0 && out_stride_b % 8 == 0 &&
                  aligned16(out),
              M4D_ERR_ALIGN);

  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* m, const void* base, int L, long long sb, long long sl, uint32_t rows) {
    const uint32_t box[4] = {64, 1, rows, 1};
    uint64_t dims[4] = {static_cast<uint64_t>(A_D), static_cast<uint64_t>(heads),
                        static_cast<uint64_t>(L), static_cast<uint64_t>(B)};
    uint64_t str[3] = {static_cast<uint64_t>(A_D) * 2, static_cast<uint64_t>(sl) * 2,
                       static_cast<uint64_t>(B > 1 ? sb : sl * L) * 2};
    return make_tmap_bf16(m, base, 4, dims, str, box);
  };
  int rc;
  const uint32_t kv_rows = A_BKV;
  if ((rc = mk(&tmQ, q, Lq, q_stride_b, q_stride_l, A_BQ)) != M4D_OK) return rc;
  if ((rc = mk(&tmK, k, Lk, kv_stride_b, kv_stride_l, kv_rows)) != M4D_OK) return rc;
  if ((rc = mk(&tmV, v, Lk, kv_stride_b, kv_stride_l, kv_rows)) != M4D_OK) return rc;

  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, AttnParams) = attn_fwd_d128_kernel<A_DEFAULT_PP, A_DEFAULT_MODE>;
  int threads = A_DEFAULT_MODE == 2 ? A_THREADS_SPLIT : A_THREADS;
  // short key ranges (the cross-attention: 769 keys = 7 tiles): persistent CTAs, one per SM
  const int n_qblk = (Lq + A_NQ * A_BQ - 1) / (A_NQ * A_BQ);
  const long long n_items = static_cast<long long>(n_qblk) * heads * B;
  bool persist = A_DEFAULT_MODE != 2 && k_lens == nullptr && Lk <= A_PERSIST_MAX_KEYS && n_items > sm_count() &&
                 n_items < (1ll << 30);
  if (persist) kern = attn_fwd_d128_kernel<A_DEFAULT_PP, A_DEFAULT_MODE == 2 ? 1 : A_DEFAULT_MODE, true>;
#ifdef M4D_DEV
  // development build only: m4d_dev_set_flags(0x100 | (MODE << 4) | PP) selects a measured variant
  if (g_dev_flags & 0x100) {
    const int pp = g_dev_flags & 0xF, mode = (g_dev_flags >> 4) & 3;
    persist = false;
    if (mode == 2) {
      kern = pp == 0 ? attn_fwd_d128_kernel<0, 2> : pp == 1 ? attn_fwd_d128_kernel<1, 2>
           : pp == 2 ? attn_fwd_d128_kernel<2, 2> : pp == 3 ? attn_fwd_d128_kernel<3, 2> : attn_fwd_d128_kernel<4, 2>;
      threads = A_THREADS_SPLIT;
    } else if (mode == 1) {
      kern = pp == 0 ? attn_fwd_d128_kernel<0, 1> : pp == 1 ? attn_fwd_d128_kernel<1, 1>
           : pp == 2 ? attn_fwd_d128_kernel<2, 1> : pp == 3 ? attn_fwd_d128_kernel<3, 1> : attn_fwd_d128_kernel<10, 1>;
      threads = A_THREADS;
    } else {
      kern = pp == 0 ? attn_fwd_d128_kernel<0, 0> : attn_fwd_d128_kernel<2, 0>;
      threads = A_THREADS;
    }
  }
#endif
  const int smem_bytes = A_SMEM_BYTES;
  rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes),
               "cudaFuncSetAttribute(attention)");
  if (rc != M4D_OK) return rc;
  AttnParams p;
  p.out = static_cast<bf16*>(out);
  p.out_stride_b = out_stride_b;
  p.out_stride_l = out_stride_l;
  p.Lq = Lq;
  p.Lk = Lk;
  p.heads = heads;
  p.k_lens = k_lens;
  p.scale_log2 = (softmax_scale > 0.f ? softmax_scale : 1.0f / sqrtf(static_cast<float>(A_D))) *
                 1.4426950408889634f;
  p.accumulate = accumulate;
  p.seg_tiles = seg_len / A_BKV;
  p.scatter_rows = n_scatter > 0 ? scatter_rows : 0;
  for (int i = 0; i < 8; ++i) p.out_scatter[i] = i < n_scatter ? static_cast<bf16*>(out_scatter[i]) : nullptr;
  p.n_qblk = n_qblk;
  p.n_items = static_cast<int>(n_items < (1ll << 30) ? n_items : 0);
  dim3 grid(n_qblk, heads, B);
  if (persist) grid = dim3(sm_count(), 1, 1);
  kern<<<grid, threads, smem_bytes, stream>>>(tmQ, tmK, tmV, p);
  M4D_CHECK_LAUNCH("attn_fwd_d128_kernel");
  return M4D_OK;
}

extern "C" int m4d_attention_fwd(const void* q, const void* k, const void* v, void* out, int B,
                                 int Lq, int Lk, int heads, int head_dim, long long q_stride_b,
                                 long long q_stride_l, long long kv_stride_b, long long kv_stride_l,
                                 long long out_stride_b, long long out_stride_l, const int* k_lens,
                                 float softmax_scale, int accumulate, void* stream_) {
  return attention_impl(q, k, v, out, B, Lq, Lk, heads, head_dim, q_stride_b, q_stride_l, kv_stride_b,
                        kv_stride_l, out_stride_b, out_stride_l, k_lens, softmax_scale, accumulate, nullptr,
                        0, 0, 0, stream_);
}

extern "C" int m4d_attention_fwd_seg2(const void* q, const void* k, const void* v, void* out, int B, int Lq,
                                      int Lk, int seg_len, int heads, int head_dim, long long q_stride_b,
                                      long long q_stride_l, long long kv_stride_b, long long kv_stride_l,
                                      long long out_stride_b, long long out_stride_l, float softmax_scale,
                                      void* stream_) {
  M4D_REQUIRE(seg_len > 0, M4D_ERR_BAD_SHAPE);
  return attention_impl(q, k, v, out, B, Lq, Lk, heads, head_dim, q_stride_b, q_stride_l, kv_stride_b,
                        kv_stride_l, out_stride_b, out_stride_l, nullptr, softmax_scale, 0, nullptr, 0, 0,
                        seg_len, stream_);
}

extern "C" int m4d_attention_fwd_scatter(const void* q, const void* k, const void* v, void* const* out, int n_out,
                                         int rows_per_out, int B, int Lq, int Lk, int heads, int head_dim,
                                         long long q_stride_b, long long q_stride_l, long long kv_stride_b,
                                         long long kv_stride_l, long long out_stride_b, long long out_stride_l,
                                         const int* k_lens, float softmax_scale, void* stream_) {
  return attention_impl(q, k, v, nullptr, B, Lq, Lk, heads, head_dim, q_stride_b, q_stride_l, kv_stride_b,
                        kv_stride_l, out_stride_b, out_stride_l, k_lens, softmax_scale, 0, out, n_out,
                        rows_per_out, 0, stream_);
}

