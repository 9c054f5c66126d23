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
constexpr int A_THREADS = 384;
constexpr int A_SMEM_BYTES = (A_NQ + A_KS + A_VS) * A_TILE_BYTES + 256 + 1024;

constexpr float A_RESCALE_THRESHOLD = 8.0f;   // log2 units

struct AttnParams {
  bf16* out;
  long long out_stride_b, out_stride_l;   // elements
  int Lq, Lk, heads;
  const int* k_lens;                      // [B] or null
  float scale_log2;                       // softmax_scale * log2(e)
  int accumulate;                         // out = bf16(out + bf16(o))  (summed cross-attention)
  int flags;                              // debug variants, see m4d_set_debug_flags
};

__global__ void __launch_bounds__(A_THREADS, 1)
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
  uint64_t* p_ready = s_full + 2;       // 2
  uint64_t* o_final = p_ready + 2;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;

  int kv_len = p.Lk;
  if (p.k_lens != nullptr) {
    int kl = p.k_lens[b];
    kv_len = kl < kv_len ? kl : kv_len;
  }
  if (kv_len < 1) kv_len = 1;
  const int n_kv = (kv_len + A_BKV - 1) / A_BKV;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9 && lane == 0) {
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
      mbar_init(&p_ready[t], 128);
      mbar_init(&o_final[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 10) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    // ------------------------------------------------------------------ data movement + MMA
    reg_dec<80>();
    if (warp == 8 && lane == 0) {
      mbar_arrive_expect_tx(q_full, A_NQ * A_TILE_BYTES);
      for (int t = 0; t < A_NQ; ++t)
        for (int h = 0; h < 2; ++h)
          tma_load_4d(sQ + t * A_TILE_BYTES + h * A_HALF_BYTES, &tmQ, q_full, h * 64, head,
                      q_blk * (A_NQ * A_BQ) + t * A_BQ, b);
      for (int j = 0; j < n_kv; ++j) {
        const int sk = j % A_KS, sv = j % A_VS;
        mbar_wait(&k_empty[sk], ((j / A_KS) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[sk], A_TILE_BYTES);
        for (int h = 0; h < 2; ++h)
          tma_load_4d(sK + sk * A_TILE_BYTES + h * A_HALF_BYTES, &tmK, &k_full[sk], h * 64, head,
                      j * A_BKV, b);
        mbar_wait(&v_empty[sv], ((j / A_VS) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[sv], A_TILE_BYTES);
        for (int h = 0; h < 2; ++h)
          tma_load_4d(sV + sv * A_TILE_BYTES + h * A_HALF_BYTES, &tmV, &v_full[sv], h * 64, head,
                      j * A_BKV, b);
      }
    } else if (warp == 9 && lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);   // V is MN-major (d contiguous)
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t tS[2] = {tmem_base + 0, tmem_base + 128};
      const uint32_t tO[2] = {tmem_base + 256, tmem_base + 384};
      // V descriptor strides: 64-wide d chunks are 16 KB apart, 8-key groups 1 KB apart.
      const uint32_t v_lbo = (p.flags & 1) ? 1024 : A_HALF_BYTES;
      const uint32_t v_sbo = (p.flags & 1) ? A_HALF_BYTES : 1024;

      auto issue_qk = [&](int t, int sk) {
#pragma unroll
        for (int kk = 0; kk < A_D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * A_HALF_BYTES + (kk & 3) * 32;
          const uint64_t ad = umma_smem_desc(q_addr + t * A_TILE_BYTES + off, 16, 1024);
          const uint64_t bd = umma_smem_desc(k_addr + sk * A_TILE_BYTES + off, 16, 1024);
          umma_ss(tS[t], ad, bd, idesc_qk, kk != 0);
        }
      };
      auto issue_pv = [&](int t, int sv, bool first) {
#pragma unroll
        for (int kk = 0; kk < A_BKV / 16; ++kk) {
          const uint64_t bd = umma_smem_desc(v_addr + sv * A_TILE_BYTES + kk * 2048, v_lbo, v_sbo);
          umma_ts(tO[t], tS[t] + kk * 8, bd, idesc_pv, !(first && kk == 0));
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      umma_commit(&s_full[0]);
      issue_qk(1, 0);
      umma_commit(&s_full[1]);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int sv = j % A_VS;
        const bool last = (j + 1 == n_kv);
        const int sk = (j + 1) % A_KS;
        mbar_wait(&v_full[sv], (j / A_VS) & 1);
        mbar_wait(&p_ready[0], j & 1);
        tc_fence_after();
        issue_pv(0, sv, j == 0);
        if (last) {
          umma_commit(&o_final[0]);
        } else {
          mbar_wait(&k_full[sk], ((j + 1) / A_KS) & 1);
          tc_fence_after();
          issue_qk(0, sk);
          umma_commit(&s_full[0]);
        }
        mbar_wait(&p_ready[1], j & 1);
        tc_fence_after();
        issue_pv(1, sv, j == 0);
        umma_commit(&v_empty[sv]);
        if (last) {
          umma_commit(&o_final[1]);
        } else {
          issue_qk(1, sk);
          umma_commit(&s_full[1]);
          umma_commit(&k_empty[sk]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    reg_inc<208>();
    const int t = warp >> 2;                       // query tile
    const int quad = warp & 3;
    const int row = quad * 32 + lane;              // row in tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + t * 128;
    const uint32_t tO = tmem_base + lane_base + 256 + t * 128;
    const float c = p.scale_log2;
    float m_used = -INFINITY;
    float l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[128];
      tmem_ld32(tS + 0, s + 0);
      tmem_ld32(tS + 32, s + 32);
      tmem_ld32(tS + 64, s + 64);
      tmem_ld32(tS + 96, s + 96);
      tmem_ld_wait();

      const int valid = kv_len - j * A_BKV;        // keys of this tile that exist
      if (valid < A_BKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) s[i] = 0xFF800000u;      // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
      float mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (j == 0) {
        m_used = mx;
      } else {
        const float m_new = fmaxf(m_used, mx);
        const bool grow = (m_new - m_used) * c > A_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, grow)) {
          const float f = fast_exp2((m_used - m_new) * c);
          m_used = m_new;
          l_sum *= f;
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
        }
      }
      const float neg_mc = -m_used * c;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[h * 64 + 2 * i + 0]), c, neg_mc));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[h * 64 + 2 * i + 1]), c, neg_mc));
          const float p2 = fast_exp2(fmaf(__uint_as_float(s[h * 64 + 2 * i + 2]), c, neg_mc));
          const float p3 = fast_exp2(fmaf(__uint_as_float(s[h * 64 + 2 * i + 3]), c, neg_mc));
          l0 += p0;
          l1 += p1;
          l2 += p2;
          l3 += p3;
          pk[i] = (p.flags & 2) ? pack_bf16(p1, p0) : pack_bf16(p0, p1);
          pk[i + 1] = (p.flags & 2) ? pack_bf16(p3, p2) : pack_bf16(p2, p3);
        }
        tmem_st32(tS + h * 32, pk);
      }
      l_sum += (l0 + l1) + (l2 + l3);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[t]);
    }

    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(&o_final[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    const int q_row = q_blk * (A_NQ * A_BQ) + t * A_BQ + row;
    const bool row_ok = q_row < p.Lq;
    bf16* orow = p.out + static_cast<long long>(b) * p.out_stride_b +
                 static_cast<long long>(q_row) * p.out_stride_l + head * A_D;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t o[32];
      tmem_ld32(tO + cc * 32, o);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem_base);
}

}  // namespace m4d

using namespace m4d;

extern "C" int m4d_attention_fwd(const void* q, const void* k, const void* v, void* out, int B,
                                 int Lq, int Lk, int heads, int head_dim, long long q_stride_b,
                                 long long q_stride_l, long long kv_stride_b, long long kv_stride_l,
                                 long long out_stride_b, long long out_stride_l, const int* k_lens,
                                 float softmax_scale, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(q && k && v && out, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(B > 0 && Lq > 0 && Lk > 0 && heads > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(head_dim == A_D, M4D_ERR_UNSUPPORTED);          // d = 128 is the Wan2.1 invariant
  M4D_REQUIRE(heads <= 65535 && B <= 65535, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(q_stride_l >= heads * A_D && kv_stride_l >= heads * A_D &&
                  out_stride_l >= heads * A_D,
              M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(q_stride_l % 8 == 0 && kv_stride_l % 8 == 0 && out_stride_l % 8 == 0 &&
                  q_stride_b % 8 == 0 && kv_stride_b % 8 == 0 && out_stride_b % 8 == 0 &&
                  aligned16(out),
              M4D_ERR_ALIGN);

  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[4] = {64, 1, 128, 1};
  auto mk = [&](CUtensorMap* m, const void* base, int L, long long sb, long long sl) {
    uint64_t dims[4] = {static_cast<uint64_t>(A_D), static_cast<uint64_t>(heads),
                        static_cast<uint64_t>(L), static_cast<uint64_t>(B)};
    uint64_t str[3] = {static_cast<uint64_t>(A_D) * 2, static_cast<uint64_t>(sl) * 2,
                       static_cast<uint64_t>(B > 1 ? sb : sl * L) * 2};
    return make_tmap_bf16(m, base, 4, dims, str, box);
  };
  int rc;
  if ((rc = mk(&tmQ, q, Lq, q_stride_b, q_stride_l)) != M4D_OK) return rc;
  if ((rc = mk(&tmK, k, Lk, kv_stride_b, kv_stride_l)) != M4D_OK) return rc;
  if ((rc = mk(&tmV, v, Lk, kv_stride_b, kv_stride_l)) != M4D_OK) return rc;

  static bool configured = false;
  if (!configured) {
    rc = cuda_ok(cudaFuncSetAttribute(attn_fwd_d128_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM_BYTES),
                 "cudaFuncSetAttribute(attention)");
    if (rc != M4D_OK) return rc;
    configured = true;
  }
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
  p.flags = g_debug_flags;
  dim3 grid((Lq + A_NQ * A_BQ - 1) / (A_NQ * A_BQ), heads, B);
  attn_fwd_d128_kernel<<<grid, A_THREADS, A_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  M4D_CHECK_LAUNCH("attn_fwd_d128_kernel");
  return M4D_OK;
}
