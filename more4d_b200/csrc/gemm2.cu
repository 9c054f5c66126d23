// 2-CTA ("cta_group::2") variant of the bf16 GEMM in gemm.cu: a CLUSTER of two CTAs on the two
// SMs of a TPC computes one 256 x 256 output tile with tcgen05.mma M = 256.
//
// Why: with one CTA per tile every MMA streams a 128x16 A slice AND a 256x16 W slice from shared
// memory, and every CTA pulls its own copy of the W tile through L2.  In a CTA pair each CTA
// holds only HALF of the W tile (128 of the 256 output columns) — the tensor cores of both SMs
// read both halves — so per-SM shared-memory traffic for W and the L2->SM traffic for W halve,
// a stage shrinks from 48 KB to 32 KB (6 stages instead of 4), and one elected thread issues
// half as many MMAs.
//
// Protocol (per pipeline stage; "leader" = cluster rank 0):
//   producer warp of EACH CTA : waits its own `empty`, then loads its 128x64 A rows and its
//       128x64 half of W with cp.async.bulk.tensor...cta_group::2, whose complete_tx lands on
//       the LEADER's `full` barrier; the leader alone arrives on it, arming it with the bytes
//       of BOTH CTAs (a remote arrival by the peer per stage halved the throughput: 775 vs
//       1480 TF/s, profiles/gemm2_r01.log).
//   MMA warp of the leader    : waits `full`, issues 4 x tcgen05.mma.cta_group::2 (M256 N256
//       K16), then tcgen05.commit...multicast::cluster releases `empty` in BOTH CTAs.
//   epilogue warpgroups of EACH CTA drain their own 128 TMEM lanes; `tmem_full` arrives by
//       multicast commit, `tmem_empty` lives in the leader and counts one arrival per epilogue
//       warp of both CTAs.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace m4d {

constexpr int G2_M = 128;            // rows per CTA (256 per cluster tile)
constexpr int G2_NH = 128;           // W rows (output columns) held per CTA
constexpr int G2_N = 256;            // output columns per tile
constexpr int G2_K = 64;
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = G2_M * G2_K * 2;     // 16 KB
constexpr int G2_B_BYTES = G2_NH * G2_K * 2;    // 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 256 + 1024;
constexpr int G2_THREADS = 384;

struct Gemm2Epi {
  const bf16* bias;
  void* out;
  long long ldo;
  const float* res;
  long long ldr;
  const float* gate;
  long long gate_bstride;
  int rows_per_batch;
};

__device__ __forceinline__ float g2_gelu_tanh(float x) {
  // 0.5 x (1 + tanh(u)) = x / (1 + exp(-2u)), u = sqrt(2/pi) (x + 0.044715 x^3): one ex2 and one
  // rcp instead of tanhf's range-split evaluation (~25 instructions); fp32-accurate to ~1e-6
  // relative, far below the bf16 rounding of the output.  exp overflow -> x / inf = -0 for very
  // negative x, exp underflow -> x for very positive x: the right limits.
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  return __fdividef(x, 1.0f + fast_exp2(-2.8853900817779268f * u));     // 2 * log2(e)
}

// Epilogue for one 32-column chunk; same arithmetic as gemm.cu (bf16-rounded Linear output).
template <int EPI>
__device__ __forceinline__ void g2_epilogue_chunk(const uint32_t* r, const Gemm2Epi& ep, long long m,
                                                  int n0, int N, bool row_ok) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  const bool full = (n0 + 32 <= N);
  if (ep.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (full || n0 + i < N) v[i] += __bfloat162float(__ldg(ep.bias + n0 + i));
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
  if (EPI == M4D_EPI_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = g2_gelu_tanh(v[i]);
  }
  if (!row_ok) return;
  if (EPI == M4D_EPI_GATE_RESIDUAL_F32) {
    float* orow = reinterpret_cast<float*>(ep.out) + m * ep.ldo + n0;
    const float* rrow = ep.res + m * ep.ldr + n0;
    const float* grow = ep.gate ? ep.gate + (m / ep.rows_per_batch) * ep.gate_bstride + n0 : nullptr;
    if (full && ((reinterpret_cast<uintptr_t>(orow) | reinterpret_cast<uintptr_t>(rrow)) & 31) == 0) {
      // 256-bit accesses: each covers a whole 32-byte sector of the fp32 row (16-byte pieces leave
      // L1 as half-sector partial writes: twice the requests and twice the L1->L2 bytes)
      float rr[32];
#pragma unroll
      for (int q = 0; q < 4; ++q) ldg256_f(rrow + q * 8, rr + q * 8);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 g = grow ? __ldg(reinterpret_cast<const float4*>(grow + q * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
        rr[q * 4 + 0] += v[q * 4 + 0] * g.x;
        rr[q * 4 + 1] += v[q * 4 + 1] * g.y;
        rr[q * 4 + 2] += v[q * 4 + 2] * g.z;
        rr[q * 4 + 3] += v[q * 4 + 3] * g.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) stg256_f(orow + q * 8, rr + q * 8);
    } else if (full) {
      float4 rr[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) rr[q] = *reinterpret_cast<const float4*>(rrow + q * 4);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 g = grow ? __ldg(reinterpret_cast<const float4*>(grow + q * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
        rr[q].x += v[q * 4 + 0] * g.x;
        rr[q].y += v[q * 4 + 1] * g.y;
        rr[q].z += v[q * 4 + 2] * g.z;
        rr[q].w += v[q * 4 + 3] * g.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(orow + q * 4) = rr[q];
    } else {
      for (int i = 0; i < 32; ++i)
        if (n0 + i < N) orow[i] = rrow[i] + v[i] * (grow ? grow[i] : 1.f);
    }
  } else {
    bf16* orow = reinterpret_cast<bf16*>(ep.out) + m * ep.ldo + n0;
    if (full && ((reinterpret_cast<uintptr_t>(orow) & 31) == 0)) {
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
      stg256(orow, o);
      stg256(orow + 16, o + 8);
    } else if (full && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
        o.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
        o.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
        o.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
        *reinterpret_cast<uint4*>(orow + q * 8) = o;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (n0 + i < N) orow[i] = __float2bfloat16_rn(v[i]);
    }
  }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     int M, int N, int K, int pw, Gemm2Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
  uint64_t* empty = full + G2_STAGES;
  uint64_t* tfull = empty + G2_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full[s], 1);          // the leader's expect_tx arrival; the peer only adds tx bytes
      mbar_init(&empty[s], 1);         // multicast commit from the leader's MMA warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * 8);    // one arrival per epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + 2 * G2_M - 1) / (2 * G2_M);      // 256-row tiles
  const int num_n = (N + G2_N - 1) / G2_N;
  const int tiles = num_m * num_n;
  const int kblocks = (K + G2_K - 1) / G2_K;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  auto tile_coords2 = [&](int tile, int& m_blk, int& n_blk) {
    const int per_panel = num_m * pw;
    int panel = tile / per_panel;
    const int full_panels = num_n / pw;
    if (panel > full_panels) panel = full_panels;
    const int n0 = panel * pw;
    const int w = min(pw, num_n - n0);
    const int local = tile - panel * per_panel;
    m_blk = local / w;
    n_blk = n0 + local - m_blk * w;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_coords2(tile, m_blk, n_blk);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_s = smem + stage * G2_STAGE_BYTES;
          uint8_t* b_s = a_s + G2_A_BYTES;
          const uint32_t full_leader = mapa_u32(&full[stage], 0);
          // The peer never arrives on `full`: its bytes are covered by the leader's expect_tx.
          // Safe because the peer's loads for (stage, phase p+1) are only issued after the
          // multicast commit that follows the MMAs of (stage, phase p), i.e. after the
          // leader's barrier has already completed phase p.
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * G2_STAGE_BYTES);
          tma_load_2d_2sm(a_s, &tmA, full_leader, kb * G2_K, m_blk * 2 * G2_M + rank * G2_M);
          tma_load_2d_2sm(b_s, &tmB, full_leader, kb * G2_K, n_blk * G2_N + rank * G2_NH);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, G2_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < tiles; tile += num_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_N;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t b_addr = a_addr + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < G2_K / 16; ++k) {
            const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_ss_2sm(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit_2sm(&empty[stage]);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_coords2(tile, m_blk, n_blk);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const long long m = static_cast<long long>(m_blk) * 2 * G2_M + rank * G2_M + quad * 32 + lane;
      const bool row_ok = m < M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * G2_N;
#pragma unroll 1
      for (int c = half * 4; c < half * 4 + 4; ++c) {
        const int n0 = n_blk * G2_N + c * 32;
        if (n0 >= N) break;
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        g2_epilogue_chunk<EPI>(r, ep, m, n0, N, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // relaxed: only the TMEM reads (complete after tcgen05.wait::ld) must precede the hand-back;
        // a release at cluster scope would wait for the epilogue's global stores (MEMBAR.ALL.GPU)
        if (leader) mbar_arrive_relaxed(&tempty[acc]);
        else mbar_arrive_cluster_relaxed(mapa_u32(&tempty[acc], 0));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync_all();                   // no CTA may exit while its peer can still touch its smem / TMEM
  if (warp == 2) tmem_dealloc_2sm<512>(tmem_base);
}

template <int EPI>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                        const Gemm2Epi& ep, cudaStream_t stream) {
  auto kern = gemm2_bf16_tn_kernel<EPI>;
  int rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES),
                   "cudaFuncSetAttribute(gemm2)");
  if (rc != M4D_OK) return rc;
  const int num_n = (N + G2_N - 1) / G2_N;
  const int tiles = ((M + 2 * G2_M - 1) / (2 * G2_M)) * num_n;
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  const long long slab = static_cast<long long>(G2_N) * K * 2;
  int pw_max = static_cast<int>((48ll << 20) / (slab > 0 ? slab : 1));
  if (pw_max < 1) pw_max = 1;
  const int panels = (num_n + pw_max - 1) / pw_max;
  const int pw = (num_n + panels - 1) / panels;
  kern<<<clusters * 2, G2_THREADS, G2_SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, pw, ep);
  M4D_CHECK_LAUNCH("gemm2_bf16_tn_kernel");
  return M4D_OK;
}

// Called from m4d_gemm_bf16 (gemm.cu) for the epilogues that have a 2-CTA instantiation.
int gemm2_dispatch(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out,
                   long long ldo, int M, int N, int K, int epilogue, const float* residual, long long ldr,
                   const float* gate, long long gate_batch_stride, int rows_per_batch,
                   cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t str[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {G2_K, G2_M};
    int rc = make_tmap_bf16(&tmA, a, 2, dims, str, box);
    if (rc != M4D_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {G2_K, G2_NH};
    int rc = make_tmap_bf16(&tmB, w, 2, dims, str, box);
    if (rc != M4D_OK) return rc;
  }
  Gemm2Epi ep;
  ep.bias = static_cast<const bf16*>(bias);
  ep.out = out;
  ep.ldo = ldo;
  ep.res = residual;
  ep.ldr = ldr;
  ep.gate = gate;
  ep.gate_bstride = gate_batch_stride;
  ep.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  switch (epilogue) {
    case M4D_EPI_BF16: return launch_gemm2<M4D_EPI_BF16>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_GELU_TANH: return launch_gemm2<M4D_EPI_GELU_TANH>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_GATE_RESIDUAL_F32:
      return launch_gemm2<M4D_EPI_GATE_RESIDUAL_F32>(tmA, tmB, M, N, K, ep, stream);
  }
  return M4D_ERR_UNSUPPORTED;
}

}  // namespace m4d
