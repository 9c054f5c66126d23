// bf16 GEMM on the 5th-gen tensor cores:  D[M,N] = A[M,K] . W[N,K]^T  (+ fused epilogue)
//
// This is nn.Linear / the (1,2,2) patch-embedding conv of the DiT (reference call sites:
// MoRe4D/models/wan_transformer4d.py:446-448,465 q/k/v/o; :481-483,527-531,553 cross-attn;
// :620-622 FFN; :720 head; :898-902 embeddings).  Both operands are K-major ("TN"), exactly
// how torch stores activations [M,K] and Linear weights [N,K], so parameters are consumed in
// place.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0   TMA producer   : 128x64 A tile + 256x64 W tile per stage, 4-stage mbarrier ring
//   warp 1   MMA issuer     : one elected thread issues tcgen05.mma (M128 N256 K16) into TMEM
//   warp 2   TMEM allocator : 512 columns = two 128x256 fp32 accumulators (double-buffered)
//   warps 4-11 epilogue     : two warpgroups, each drains one 128-column half of the
//                             accumulator: tcgen05.ld -> fused epilogue -> global
// so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Epilogues reproduce the reference's autocast rounding: the Linear result is rounded to bf16
// before anything else touches it (oracle/dit_oracle.py Arith.linear).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace m4d {

constexpr int GM = 128, GN = 256, GK = 64, GSTAGES = 4;
constexpr int G_A_BYTES = GM * GK * 2;            // 16 KB
constexpr int G_B_BYTES = GN * GK * 2;            // 32 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_SMEM_BYTES = GSTAGES * G_STAGE_BYTES + 256 + 1024;  // + barriers + align slack
constexpr int G_THREADS = 384;               // 4 control warps + 2 epilogue warpgroups

struct GemmEpi {
  const bf16* bias;      // [N] or null
  void* out;             // bf16 or fp32, row stride ldo (elements)
  long long ldo;
  const float* res;      // fp32 residual, row stride ldr (may alias out)
  long long ldr;
  const float* gate;     // fp32 [batch, N] with batch stride gate_bstride, or null (= 1)
  long long gate_bstride;
  int rows_per_batch;    // rows of A per gate batch entry
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(u)) = x / (1 + exp(-2u)), u = sqrt(2/pi) (x + 0.044715 x^3): one ex2 and one
  // rcp instead of tanhf's range-split evaluation (~25 instructions); fp32-accurate to ~1e-6
  // relative, far below the bf16 rounding of the output.  exp overflow -> x / inf = -0 for very
  // negative x, exp underflow -> x for very positive x: the right limits.
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  return __fdividef(x, 1.0f + fast_exp2(-2.8853900817779268f * u));     // 2 * log2(e)
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f));
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t* r, const GemmEpi& ep, long long m,
                                               int n0, int N, bool row_ok) {
  // r: 32 fp32 accumulators for columns n0..n0+31 of row m
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  const bool full = (n0 + 32 <= N);
  if (EPI != M4D_EPI_F32_RAW && ep.bias != nullptr) {
    if (full) {
      const uint4* bp = reinterpret_cast<const uint4*>(ep.bias + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 b = __ldg(bp + q);
        const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[q * 8 + e * 2] += __uint_as_float(w[e] << 16);
          v[q * 8 + e * 2 + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < N) v[i] += __bfloat162float(ep.bias[n0 + i]);
    }
  }
  // the Linear output is a bf16 tensor on the reference path
  if (EPI != M4D_EPI_F32_RAW) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
  }
  if (EPI == M4D_EPI_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
  } else if (EPI == M4D_EPI_GELU_ERF) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (!row_ok) return;

  if (EPI == M4D_EPI_GATE_RESIDUAL_F32 || EPI == M4D_EPI_F32 || EPI == M4D_EPI_F32_RAW) {
    float* orow = reinterpret_cast<float*>(ep.out) + m * ep.ldo + n0;
    if (EPI == M4D_EPI_GATE_RESIDUAL_F32) {
      const float* rrow = ep.res + m * ep.ldr + n0;
      const float* grow =
          ep.gate ? ep.gate + (m / ep.rows_per_batch) * ep.gate_bstride + n0 : nullptr;
      if (full) {
        // all residual loads first: `out` may alias `residual` (x is updated in place), so the
        // compiler must not be left to interleave dependent load/store round trips
        float4 rr[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) rr[q] = *reinterpret_cast<const float4*>(rrow + q * 4);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 g = grow ? __ldg(reinterpret_cast<const float4*>(grow + q * 4))
                          : make_float4(1.f, 1.f, 1.f, 1.f);
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
      if (full && (reinterpret_cast<uintptr_t>(orow) & 31) == 0) {
        // 256-bit stores cover whole 32-byte sectors (16-byte pieces leave L1 as partial writes)
#pragma unroll
        for (int q = 0; q < 4; ++q) stg256_f(orow + q * 8, v + q * 8);
      } else if (full) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(orow + q * 4) =
              make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      } else {
        for (int i = 0; i < 32; ++i)
          if (n0 + i < N) orow[i] = v[i];
      }
    }
  } else {
    bf16* orow = reinterpret_cast<bf16*>(ep.out) + m * ep.ldo + n0;
    if (EPI == M4D_EPI_ADD_BF16) {
      const bf16* rrow = reinterpret_cast<const bf16*>(ep.res) + m * ep.ldr + n0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < N) v[i] += __bfloat162float(rrow[i]);
    }
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

// Tile rasterisation: N is cut into panels of `pw` tile-columns whose weight slab
// (pw * 256 * K * 2 B) fits the L2 next to the streaming A operand; inside a panel tiles run
// n-fastest so the CTAs in flight share both their A rows and the panel's weights.  Without
// panels a K = 13824 / N = 13824 weight matrix (141 MB > 126 MB L2) is re-streamed from HBM by
// every wave of CTAs.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int pw, int& m_blk,
                                            int& n_blk) {
  const int per_panel = num_m * pw;
  int panel = tile / per_panel;
  const int full_panels = num_n / pw;            // the last panel may be narrower
  if (panel > full_panels) panel = full_panels;
  const int n0 = panel * pw;
  const int w = min(pw, num_n - n0);
  const int local = tile - panel * per_panel;
  m_blk = local / w;
  n_blk = n0 + local - m_blk * w;
}

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    int M, int N, int K, int pw, GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GSTAGES * G_STAGE_BYTES);
  uint64_t* empty = full + GSTAGES;
  uint64_t* tfull = empty + GSTAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (M + GM - 1) / GM;
  const int num_n = (N + GN - 1) / GN;
  const int tiles = num_m * num_n;
  const int kblocks = (K + GK - 1) / GK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, pw, m_blk, n_blk);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_s = smem + stage * G_STAGE_BYTES;
          uint8_t* b_s = a_s + G_A_BYTES;
          mbar_arrive_expect_tx(&full[stage], G_STAGE_BYTES);
          tma_load_2d(a_s, &tmA, &full[stage], kb * GK, m_blk * GM);
          tma_load_2d(b_s, &tmB, &full[stage], kb * GK, n_blk * GN);
          if (++stage == GSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GM, GN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * G_STAGE_BYTES);
          const uint32_t b_addr = a_addr + G_A_BYTES;
#pragma unroll
          for (int k = 0; k < GK / 16; ++k) {
            const uint64_t ad = umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == GSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;              // which 128-column half this warpgroup drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, pw, m_blk, n_blk);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const long long m = static_cast<long long>(m_blk) * GM + quad * 32 + lane;
      const bool row_ok = m < M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * GN;
      if (EPI == M4D_EPI_GATE_RESIDUAL_F32) {
        // warm the L2->L1 path of this row's residual (256 fp32 = 8 lines) before the first
        // dependent load: under HBM load the per-chunk load latency is otherwise serialised
        // 8 times per tile and the epilogue, not the MMA main loop, paces the kernel.
        if (row_ok) {
          const float* rrow = ep.res + m * ep.ldr + static_cast<long long>(n_blk) * GN;
#pragma unroll
          for (int c = half * 4; c < half * 4 + 4; ++c)
            if (n_blk * GN + c * 32 < N)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(rrow + c * 32));
        }
      }
#pragma unroll 1
      for (int c = half * 4; c < half * 4 + 4; ++c) {
        const int n0 = n_blk * GN + c * 32;
        if (n0 >= N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_row + c * 32, r);
        tmem_ld_wait();
        epilogue_chunk<EPI>(r, ep, m, n0, N, row_ok);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

template <int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                       const GemmEpi& ep, cudaStream_t stream) {
  auto kern = gemm_bf16_tn_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    int rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          G_SMEM_BYTES),
                     "cudaFuncSetAttribute(gemm)");
    if (rc != M4D_OK) return rc;
    configured = true;
  }
  const int tiles = ((M + GM - 1) / GM) * ((N + GN - 1) / GN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  // panel width: weight slab of a panel <= ~48 MB, panels of (nearly) equal width
  const int num_n = (N + GN - 1) / GN;
  const long long slab = static_cast<long long>(GN) * K * 2;
  constexpr long long budget = 48ll << 20;        // 32 / 48 / 72 MB measured within noise (profiles/gemm_r01.md)
  int pw_max = static_cast<int>(budget / (slab > 0 ? slab : 1));
  if (pw_max < 1) pw_max = 1;
  const int panels = (num_n + pw_max - 1) / pw_max;
  const int pw = (num_n + panels - 1) / panels;
  kern<<<grid, G_THREADS, G_SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, pw, ep);
  M4D_CHECK_LAUNCH("gemm_bf16_tn_kernel");
  return M4D_OK;
}

}  // namespace m4d

using namespace m4d;

extern "C" int m4d_gemm_bf16(const void* a, long long lda, const void* w, long long ldw,
                             const void* bias, void* out, long long ldo, int M, int N, int K,
                             int epilogue, const void* residual_, long long ldr, const float* gate,
                             long long gate_batch_stride, int rows_per_batch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float* residual = static_cast<const float*>(residual_);
  M4D_REQUIRE(M > 0 && N > 0 && K > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(a && w && out, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, M4D_ERR_ALIGN);
  M4D_REQUIRE(lda >= K && ldw >= K && ldo >= N, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(epilogue >= 0 && epilogue < M4D_EPI_COUNT, M4D_ERR_UNSUPPORTED);
  if (epilogue == M4D_EPI_GATE_RESIDUAL_F32) {
    M4D_REQUIRE(residual != nullptr && ldr >= N, M4D_ERR_BAD_SHAPE);
    M4D_REQUIRE(rows_per_batch > 0 || gate == nullptr, M4D_ERR_BAD_SHAPE);
    M4D_REQUIRE(aligned16(residual) && ldr % 4 == 0 && (gate == nullptr || aligned16(gate)) &&
                    gate_batch_stride % 4 == 0,
                M4D_ERR_ALIGN);
  }
  if (epilogue == M4D_EPI_GATE_RESIDUAL_F32 || epilogue == M4D_EPI_F32 || epilogue == M4D_EPI_F32_RAW)
    M4D_REQUIRE(aligned16(out) && ldo % 4 == 0, M4D_ERR_ALIGN);
  if (epilogue == M4D_EPI_ADD_BF16) M4D_REQUIRE(residual != nullptr && ldr >= N, M4D_ERR_BAD_SHAPE);
  if (bias) M4D_REQUIRE(aligned16(bias), M4D_ERR_ALIGN);

  // 2-CTA kernel (gemm2.cu) for the large-M block GEMMs (development builds: flag 0x2000 forces
  // the 1-CTA kernel for A/B timing).
  {
#ifdef M4D_DEV
    const bool want = !(g_dev_flags & 0x2000);
#else
    const bool want = true;
#endif
    const bool has = epilogue == M4D_EPI_BF16 || epilogue == M4D_EPI_GELU_TANH ||
                     epilogue == M4D_EPI_GATE_RESIDUAL_F32;
    if (want && has && M >= 1024 && N >= 256)
      return gemm2_dispatch(a, lda, w, ldw, bias, out, ldo, M, N, K, epilogue, residual, ldr, gate,
                            gate_batch_stride, rows_per_batch, stream);
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t str[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {GK, GM};
    int rc = make_tmap_bf16(&tmA, a, 2, dims, str, box);
    if (rc != M4D_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {GK, GN};
    int rc = make_tmap_bf16(&tmB, w, 2, dims, str, box);
    if (rc != M4D_OK) return rc;
  }
  GemmEpi ep;
  ep.bias = static_cast<const bf16*>(bias);
  ep.out = out;
  ep.ldo = ldo;
  ep.res = residual;
  ep.ldr = ldr;
  ep.gate = gate;
  ep.gate_bstride = gate_batch_stride;
  ep.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  switch (epilogue) {
    case M4D_EPI_BF16: return launch_gemm<M4D_EPI_BF16>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_GELU_TANH: return launch_gemm<M4D_EPI_GELU_TANH>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_GELU_ERF: return launch_gemm<M4D_EPI_GELU_ERF>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_F32: return launch_gemm<M4D_EPI_F32>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_GATE_RESIDUAL_F32:
      return launch_gemm<M4D_EPI_GATE_RESIDUAL_F32>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_ADD_BF16: return launch_gemm<M4D_EPI_ADD_BF16>(tmA, tmB, M, N, K, ep, stream);
    case M4D_EPI_F32_RAW: return launch_gemm<M4D_EPI_F32_RAW>(tmA, tmB, M, N, K, ep, stream);
  }
  return M4D_ERR_UNSUPPORTED;
}
