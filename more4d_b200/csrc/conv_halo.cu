// 3x3 (x kt) stride-1 convolution over channels-last activations with the input HALO staged once
// in shared memory — the kernel behind CausalConv3d 3x3x3 (MoRe4D/models/wan_vae.py:21-40), the
// 3x3 Conv2d's of Resample's upsample branch (:82-92) and of the adaptors' ResnetBlocks
// (MoRe4D/models/trajectory_module.py:73-87): > 95 % of the Motion-Sensitive VAE's FLOPs.
//
// Why a second kernel: conv.cu issues one TMA box per tap, so every output tile pulls its input
// window 27 times and its weights once per 128 pixels through L2; ncu shows it bound by the
// L2->SM fabric (11.5 TB/s of TMA traffic, tensor pipe 43-47 %, profiles/conv_r01.md).  Here
//   * one CTA tile is 16 x 16 output pixels of one frame = TWO M=128 sub-tiles that share every
//     weight stage (weights stream once per 256 pixels), and
//   * per time tap and 64-channel block ONE TMA box {64 ch, 18 w, 18 h} lands the haloed input
//     window as [18][18] pixels x 128 B, SWIZZLE_128B; the nine spatial taps are nine UMMA
//     descriptors into that same buffer: GEMM row m = (hh, ww) of a sub-tile is pixel
//     (hh + b, ww + c + 8 sub) of the window, i.e. descriptor start = tap shift, 8-row groups =
//     8 consecutive pixels of a row (128 B apart), stride between groups = one window row
//     (18 * 128 B).  The 128-byte swizzle is a function of the shared-memory ADDRESS bits for
//     both TMA and tcgen05.mma, so shifted starts read back exactly what TMA wrote.
// L2->SM traffic per output pixel drops ~3.3x for 96->96 channels (1134 KB / 128 px ->
// 673 KB / 256 px) and more for wider layers.  Zero padding in space and the causal padding in
// time are TMA out-of-bounds zero fill (negative box coordinates), as in conv.cu.
//
// CTA = 384 threads, persistent: warp 0 halo producer, warp 3 weight producer, warp 1 MMA
// issuer, warp 2 TMEM allocator, warps 4-7 / 8-11 epilogue of sub-tile 0 / 1.  Accumulators:
// 2 sub-tiles x NT columns, double-buffered when 4 * NT <= 512.
#include "conv_common.cuh"

namespace m4d {

constexpr int CH_T = 16;                         // output tile is CH_T x CH_T pixels
constexpr int CH_HW = CH_T + 2;                  // haloed window edge
constexpr int CH_A_BYTES = CH_HW * CH_HW * 128;  // 41472: [18][18] pixels x 64 channels
constexpr int CH_A_STRIDE = 41 * 1024;           // ring stride (1024-B aligned for SWIZZLE_128B)
constexpr int CH_ROW_BYTES = CH_HW * 128;        // one window row = stride between 8-pixel groups
constexpr int CH_THREADS = 384;
constexpr int CH_MAX_A = 3, CH_MAX_B = 9;
constexpr int CH_SMEM_MAX = 227 * 1024;

// SWIZZLE_128B K-major descriptors as {lo, hi}: lo = (address >> 4) | LBO(16 B) << 16, hi = SBO >> 4 |
// version 1 | layout SWIZZLE_128B.  The base-offset field stays 0 although tap-shifted starts are
// not 1024-B aligned: measured on the B200, the swizzle XOR is taken from the absolute address
// bits (a non-zero base offset gives wrong results), matching what TMA wrote.
constexpr uint32_t CH_DESC_HI_A = (CH_ROW_BYTES >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t CH_DESC_HI_B = (1024 >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ uint64_t desc_pair(uint32_t lo, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// All MMAs of one spatial tap: both sub-tiles x KS k-slices of 16 channels.
template <bool PAIR>
__device__ __forceinline__ void umma_ss_x(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR) umma_ss_2sm(d, a, b, idesc, acc);
  else umma_ss(d, a, b, idesc, acc);
}
template <bool PAIR>
__device__ __forceinline__ void umma_commit_x(uint64_t* bar) {
  if (PAIR) umma_commit_2sm(bar);          // arrives on this barrier in BOTH CTAs of the pair
  else umma_commit(bar);
}

template <int KS, bool PAIR>
__device__ __forceinline__ void issue_tap(uint32_t d0, uint32_t d1, uint32_t a_lo, uint32_t b_lo,
                                          uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int k = 0; k < KS; ++k)
    umma_ss_x<PAIR>(d0, desc_pair(a_lo + 2 * k, CH_DESC_HI_A), desc_pair(b_lo + 2 * k, CH_DESC_HI_B), idesc,
                    k == 0 ? acc0 : 1u);
#pragma unroll
  for (int k = 0; k < KS; ++k)
    umma_ss_x<PAIR>(d1, desc_pair(a_lo + 64 + 2 * k, CH_DESC_HI_A), desc_pair(b_lo + 2 * k, CH_DESC_HI_B), idesc,
                    k == 0 ? acc0 : 1u);
}

// Epilogue with the NEXT layer's RMS_norm (+ SiLU) fused in (wan_vae.py:43-58 after :198-202):
// the thread owns all NTC = Cout channels of its pixel, so the L2 norm over channels is a
// thread-local reduction.  Same rounding points as rmsnorm_silu_cl_kernel: norm rounded to
// bf16, then y / norm, * sqrt(C), * gamma, SiLU each rounded to bf16.  Writes the raw output
// (when p.out != null) and the normalised one; removes one read + one write of the activation
// and a kernel launch per RMS_norm.
//
// Two passes so that the accumulators can be handed back EARLY: pass 1 drains TMEM (bias,
// rounding, residual, raw store, sum of squares) into packed bf16 registers; the caller then
// releases the TMEM buffer and pass 2 (the normalisation: 2/3 of the epilogue's instructions)
// runs while the tensor pipe is already on the next tile.  At NT = 192 the accumulators are
// single-buffered (2 sub-tiles x 192 columns x 2 > 512), so everything before the release is
// exposed: 7.4 ms of a 23.6 ms launch at 49x360x640 before the split (profiles/conv_fused_r02.md).
// TMEM loads are double-buffered (chunk c+1 is in flight while chunk c is processed; the
// single-buffered version spent 40 % of its samples on the TMEM / residual scoreboards), all
// bf16 arithmetic is packed (ptx.cuh), and the epilogue warps run with 208 registers
// (setmaxnreg; producers / issuer 80) to hold up to 96 packed words + two TMEM chunks + two
// residual chunks.
template <int NTC, bool NORM>
__device__ __forceinline__ float drain_pass1(const ConvParams& p, uint32_t t_row, long long pix_off, int n0,
                                             bool pix_ok, bool has_res, uint4* rnext, uint32_t* yp) {
  float ss = 0.f;
  // TMEM loads double-buffered: 32-column chunks while the registers allow it (NTC <= 128: half as
  // many load round trips per tile), 16-column halves at 192 channels (96 packed words are live)
  constexpr int TW = NTC <= 128 ? 32 : 16;
  uint32_t rr[2][TW];
  if (TW == 32) tmem_ld32(t_row, rr[0]);
  else tmem_ld16(t_row, rr[0]);
#pragma unroll
  for (int c0 = 0; c0 < NTC; c0 += 32) {
    uint4 rcur[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
    if (has_res && c0 + 32 < NTC) load_res_chunk32(p.residual + pix_off + c0 + 32, rnext);   // one chunk ahead
    uint32_t* y = yp + (c0 >> 1);
    if (TW == 32) {
      const int cur = (c0 >> 5) & 1;
      tmem_ld_wait();
      if (c0 + 32 < NTC) tmem_ld32(t_row + c0 + 32, rr[cur ^ 1]);
      conv_packed<4>(rr[cur], p.bias ? p.bias + n0 + c0 : nullptr, has_res ? rcur : nullptr, y);
    } else {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int cc = c0 + hf * 16;
        const int cur = (cc >> 4) & 1;
        tmem_ld_wait();
        if (cc + 16 < NTC) tmem_ld16(t_row + cc + 16, rr[cur ^ 1]);
        conv_packed<2>(rr[cur], p.bias ? p.bias + n0 + cc : nullptr, has_res ? rcur + 2 * hf : nullptr, y + hf * 8);
      }
    }
    if (NORM) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = bf16_lo(y[i]), b = bf16_hi(y[i]);
        ss += a * a + b * b;
      }
    }
  }
  return ss;
}

template <int NTC>
__device__ __forceinline__ void rmsnorm_pass2(const ConvParams& p, long long pix_off, const uint32_t* yp, float ss) {
  const float inv = 1.0f / fmaxf(bf16_round(sqrtf(ss)), 1e-12f);
  const float sc = sqrtf(static_cast<float>(NTC));
  bf16* o = p.norm_out + pix_off;
#pragma unroll
  for (int c0 = 0; c0 < NTC; c0 += 16) {            // 16 channels = one 256-bit store
    const uint4 ga = __ldg(reinterpret_cast<const uint4*>(p.norm_gamma + c0));
    const uint4 gb = __ldg(reinterpret_cast<const uint4*>(p.norm_gamma + c0 + 8));
    const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    uint32_t ow[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ow[e] = rmsnorm_tail_bf16x2(yp[(c0 >> 1) + e], inv, sc, gw[e]);
      if (p.norm_silu) ow[e] = silu_bf16x2(ow[e]);
    }
    stg256(o + c0, ow);
  }
}

// GroupNorm(32 groups x 4 channels) statistics of one tile (trajectory_module.py:54-60): every
// thread holds its pixel's 128 bf16 outputs; per group (sum, sum of squares) -> 64 values, summed
// over the warp's 32 pixels by recursive halving (62 shuffles: lane L ends with group L's pair),
// over the CTA's 8 epilogue warps through shared memory, and stored as this tile's partial.  The
// order of every addition is fixed, so the statistics are reproducible run to run.
__device__ __forceinline__ void gn_tile_stats(const uint32_t* yp, bool pix_ok, float* gsm /*[8][64]*/,
                                              float* partial /*[64]*/, int ewarp, int lane) {
  float v[64];
#pragma unroll
  for (int g = 0; g < 32; ++g) {
    const float a = bf16_lo(yp[2 * g]), b = bf16_hi(yp[2 * g]), c = bf16_lo(yp[2 * g + 1]), d = bf16_hi(yp[2 * g + 1]);
    v[2 * g] = pix_ok ? (a + b) + (c + d) : 0.f;
    v[2 * g + 1] = pix_ok ? (a * a + b * b) + (c * c + d * d) : 0.f;
  }
#pragma unroll
  for (int half = 32, m = 16; m >= 1; half >>= 1, m >>= 1) {
    const bool upper = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  *reinterpret_cast<float2*>(gsm + ewarp * 64 + 2 * lane) = make_float2(v[0], v[1]);
  named_bar_sync(1, 256);                              // the 8 epilogue warps
  if (ewarp < 2) {
    const int j = ewarp * 32 + lane;
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += gsm[w8 * 64 + j];
    partial[j] = t;
  }
}

// One tile of the vector epilogue: pass 1, release the accumulators, pass 2 (NORM: the fused
// RMS_norm; otherwise just the stores of the full NTC-channel row).
template <int NTC, bool NORM, typename Release>
__device__ __forceinline__ void epilogue_vec(const ConvParams& p, uint32_t t_row, long long pix_off, int n0,
                                             bool pix_ok, bool has_res, uint4* rnext, Release& release, int acc,
                                             int lane, float* gsm = nullptr, float* partial = nullptr,
                                             int ewarp = 0) {
  uint32_t yp[NTC / 2];
  const float ss = drain_pass1<NTC, NORM>(p, t_row, pix_off, n0, pix_ok, has_res, rnext, yp);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) release(acc);
  if (NTC == 128 && !NORM) {
    if (partial != nullptr) gn_tile_stats(yp, pix_ok, gsm, partial, ewarp, lane);   // warp-uniform branch
  }
  if (!pix_ok) return;
  if (NORM) {
    if (p.out != nullptr) {                         // the raw output (x + h) some later shortcut consumes
#pragma unroll
      for (int c0 = 0; c0 < NTC; c0 += 32)
        store_chunk_packed32(reinterpret_cast<bf16*>(p.out) + pix_off + c0, yp + (c0 >> 1));
    }
    rmsnorm_pass2<NTC>(p, pix_off, yp, ss);
  } else {
#pragma unroll
    for (int c0 = 0; c0 < NTC; c0 += 32)
      store_chunk_packed32(reinterpret_cast<bf16*>(p.out) + pix_off + c0, yp + (c0 >> 1));
  }
}

// The nine spatial taps of one (time tap, 64-channel block): wait for each weight stage, issue,
// release it.  Ring slot and parity are compile-time functions of the tap; the next stage's
// barrier is probed before the current stage's MMAs are issued so its latency is hidden.
template <int BST, int KS, int TPB, bool PAIR>
__device__ __forceinline__ void mma_block(uint64_t* b_full, uint64_t* b_empty, uint64_t* a_empty_bar,
                                          uint32_t blk, uint32_t a_lo, uint32_t b_lo0, uint32_t b_step,
                                          uint32_t d0, uint32_t d1, uint32_t idesc, bool first_block,
                                          bool ready) {
  constexpr int G = (9 + TPB - 1) / TPB;
  constexpr int per = G / BST;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int sb = g % BST;
    if (!ready) mbar_wait(&b_full[sb], (blk * per + g / BST) & 1);
    tc_fence_after();
    if (g + 1 < G) ready = mbar_try_wait(&b_full[(g + 1) % BST], (blk * per + (g + 1) / BST) & 1);
    const uint32_t b_lo = b_lo0 + sb * b_step;
    if (elect_one()) {
      if (TPB == 1) {
        const uint32_t a_tap = a_lo + ((g / 3) * CH_HW + (g % 3)) * 8;
        issue_tap<KS, PAIR>(d0, d1, a_tap, b_lo, idesc, (first_block && g == 0) ? 0u : 1u);
      } else {
        // thin input: the box holds TPB taps x KS k-slices; slice q -> tap g*TPB + q/KS
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int tap9 = g * TPB + q / KS;
            if (tap9 < 9) {
              const uint32_t a_tap = a_lo + ((tap9 / 3) * CH_HW + (tap9 % 3) + 8 * sub) * 8 + 2 * (q % KS);
              umma_ss_x<PAIR>(sub ? d1 : d0, desc_pair(a_tap, CH_DESC_HI_A), desc_pair(b_lo + 2 * q, CH_DESC_HI_B),
                              idesc, (first_block && g == 0 && q == 0) ? 0u : 1u);
            }
          }
      }
      umma_commit_x<PAIR>(&b_empty[sb]);
      if (g == G - 1) umma_commit_x<PAIR>(a_empty_bar);
    }
  }
}

// BST = weight-ring depth, 3 or 9: a divisor of the 9 spatial taps, so ring slot and mbarrier
// parity of every tap are compile-time functions of the tap index and of one running block
// counter — the MMA issuer does no ring arithmetic between taps (see the issuer's comment).
//
// TPB > 1 is the thin-input mode (Cin = 64 / TPB in {32, 16}: the zero-padded 3-channel model
// inputs and the 16-channel latent): K is tap-major / channel-minor, so ONE 64-wide weight box
// covers TPB consecutive taps and a block of nine taps needs G = ceil(9 / TPB) weight stages
// (BST = G) instead of nine half-empty ones; k-slice q of a box belongs to tap g*TPB + q/KS.
//
// PAIR: two CTAs of a cluster (one TPC) work on two spatial tiles with ONE weight stream: every
// MMA is a tcgen05.mma.cta_group::2 with M = 256 — rows 0-127 are this sub-tile of the leader's
// window, rows 128-255 the same sub-tile of the peer's window, each read from its own CTA's
// shared memory into its own TMEM — against a weight stage of which each CTA loads and holds
// HALF the rows (NT/2).  Per SM that halves the weight bytes coming over the L2->SM fabric (two
// thirds of the 9.2 TB/s the single-CTA kernel pulled at 96 -> 96 channels, profiles/
// conv_fused_r02.md) and the B-operand shared-memory reads (the N <= 128 layers were bound by
// A 4 KB + B N*32 B per 48..64-clock MMA).  Protocol as in gemm2.cu: all TMA loads complete on the
// LEADER's full barriers (the leader arms them with both CTAs' bytes), the leader's issuer
// commits with a multicast arrival that frees the stage / publishes the accumulators in both
// CTAs, and both CTAs' epilogue warps arrive on the leader's tempty.
template <int BST, int TPB, bool PAIR>
__global__ void __cluster_dims__(PAIR ? 2 : 1, 1, 1) __launch_bounds__(CH_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_bytes = (PAIR ? p.NT / 2 : p.NT) * 128;     // this CTA's part of one weight stage
  uint8_t* sA = smem;
  uint8_t* sB = sA + p.a_stages * CH_A_STRIDE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + BST * b_bytes);
  uint64_t* a_empty = a_full + CH_MAX_A;
  uint64_t* b_full = a_empty + CH_MAX_A;
  uint64_t* b_empty = b_full + CH_MAX_B;
  uint64_t* tfull = b_empty + CH_MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* gsm = reinterpret_cast<float*>(tmem_slot + 4);    // [2][8][64] floats, GroupNorm-statistics mode only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmX);
  if (warp == 3 && lane == 0) tma_prefetch_desc(&tmW);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CH_MAX_A; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < CH_MAX_B; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 16 : 8);      // one arrival per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_2sm<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_w = (p.W_out + CH_T - 1) / CH_T;
  const int tiles_h = (p.H_out + CH_T - 1) / CH_T;
  const int n_spatial = p.T_out * tiles_h * tiles_w;
  // work items: (spatial tile [pair], n_blk), n_blk fastest.  A pair's CTAs take spatial tiles
  // 2s and 2s+1; an odd tail gives the peer a tile beyond the last frame (TMA zero-fills its
  // loads, its epilogue stores nothing).
  const int items = (PAIR ? (n_spatial + 1) / 2 : n_spatial) * p.n_tiles;
  const int item0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int item_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int cblocks = (p.Cin + 63) / 64;
  auto decode = [&](int item, int& n_blk, int& w_blk, int& h_blk, int& t) {
    n_blk = item % p.n_tiles;
    int r = item / p.n_tiles;
    if (PAIR) r = r * 2 + static_cast<int>(rank);
    w_blk = r % tiles_w; r /= tiles_w;
    h_blk = r % tiles_h;
    t = r / tiles_h;                                   // >= T_out for the odd tail's peer tile
  };

  // register split (setmaxnreg, issued inside each role's branch so that ptxas budgets the
  // branch with it): producers / issuer / allocator 80, epilogue warps 208
  if (warp == 0) {
    // ------------------------------------------------------------ halo producer
    reg_dec<80>();
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      for (int item = item0; item < items; item += item_step) {
        int n_blk, w_blk, h_blk, t;
        decode(item, n_blk, w_blk, h_blk, t);
        for (int a = 0; a < p.kt; ++a)
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(&a_empty[sa], pa ^ 1);
            if (PAIR) {
              // the peer never arrives on `a_full`: its bytes are covered by the leader's
              // expect_tx (safe: a stage is only refilled after the multicast commit that
              // followed its MMAs, i.e. after the leader's barrier completed the phase)
              if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * CH_A_BYTES);
              tma_load_4d_2sm(sA + sa * CH_A_STRIDE, &tmX, mapa_u32(&a_full[sa], 0), cb * 64,
                              w_blk * CH_T - 1, h_blk * CH_T - 1, t + a - p.pt);
            } else {
              mbar_arrive_expect_tx(&a_full[sa], CH_A_BYTES);
              tma_load_4d(sA + sa * CH_A_STRIDE, &tmX, &a_full[sa], cb * 64, w_blk * CH_T - 1,
                          h_blk * CH_T - 1, t + a - p.pt);
            }
            if (++sa == p.a_stages) {
              sa = 0;
              pa ^= 1;
            }
          }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ weight producer
    reg_dec<80>();
    if (lane == 0) {
      uint32_t blk = 0;                                // running (time tap, channel block) counter
      for (int item = item0; item < items; item += item_step) {
        const int n_blk = item % p.n_tiles;
        for (int a = 0; a < p.kt; ++a)
          for (int cb = 0; cb < cblocks; ++cb, ++blk) {
            constexpr int G = (9 + TPB - 1) / TPB;     // weight stages per block
#pragma unroll
            for (int g = 0; g < G; ++g) {
              constexpr int per = G / BST;             // uses of a ring slot per block (odd)
              const int sb = g % BST;
              const uint32_t use = blk * per + g / BST;
              mbar_wait(&b_empty[sb], (use & 1) ^ 1);
              const int k0 = (a * 9 + g * TPB) * p.Cin + cb * 64;
              if (PAIR) {
                if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * b_bytes);
                tma_load_2d_2sm(sB + sb * b_bytes, &tmW, mapa_u32(&b_full[sb], 0), k0,
                                n_blk * p.NT + static_cast<int>(rank) * (p.NT / 2));
              } else {
                mbar_arrive_expect_tx(&b_full[sb], b_bytes);
                tma_load_2d(sB + sb * b_bytes, &tmW, &b_full[sb], k0, n_blk * p.NT);
              }
            }
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // The whole warp runs the loops converged (operands stay in uniform registers); one
    // elected lane issues.  Descriptors are {lo, hi} pairs whose hi word is constant and whose
    // lo word only takes compile-time offsets per (tap, sub-tile, k-slice).  This thread's
    // serial latency per weight stage (mbarrier probe, ring arithmetic, R2UR moves, commit) is
    // what bounded the kernel (profiles/conv_halo_r01.md: tensor pipe 56 % at N = 96, 75 % at
    // N = 192 with ~250 clocks of bookkeeping per stage), hence: static ring slots / parities
    // (BST), and the NEXT stage's barrier is probed before the current stage's MMAs are issued.
    reg_dec<80>();
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, p.NT, 0, 0);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      const uint32_t b_lo0 = ((b_base >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t b_step = static_cast<uint32_t>(b_bytes) >> 4;
      int sa = 0;
      uint32_t pa = 0, blk = 0;
      int it = 0;
      for (int item = item0; item < items; item += item_step, ++it) {
        const int acc = it % p.acc_bufs;
        const uint32_t acc_phase = (it / p.acc_bufs) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 2 * p.acc_stride;
        const uint32_t d_tmem1 = d_tmem + p.acc_stride;
        for (int a = 0; a < p.kt; ++a)
          for (int cb = 0; cb < cblocks; ++cb, ++blk) {
            const int ch_left = p.Cin - cb * 64;
            const int kslices = ch_left >= 64 ? 4 : (ch_left >> 4);
            constexpr int per = ((9 + TPB - 1) / TPB) / BST;
            const bool ready = mbar_try_wait(&b_full[0], (blk * per) & 1);
            mbar_wait(&a_full[sa], pa);
            const uint32_t a_lo = (((a_base + sa * CH_A_STRIDE) >> 4) & 0x3FFF) | (1u << 16);
            const bool first = (a | cb) == 0;
            if (TPB > 1) {
              mma_block<BST, 4 / TPB, TPB, PAIR>(b_full, b_empty, &a_empty[sa], blk, a_lo, b_lo0, b_step, d_tmem, d_tmem1, idesc, first, ready);
            } else {
              switch (kslices) {
                case 4: mma_block<BST, 4, 1, PAIR>(b_full, b_empty, &a_empty[sa], blk, a_lo, b_lo0, b_step, d_tmem, d_tmem1, idesc, first, ready); break;
                case 2: mma_block<BST, 2, 1, PAIR>(b_full, b_empty, &a_empty[sa], blk, a_lo, b_lo0, b_step, d_tmem, d_tmem1, idesc, first, ready); break;
                case 3: mma_block<BST, 3, 1, PAIR>(b_full, b_empty, &a_empty[sa], blk, a_lo, b_lo0, b_step, d_tmem, d_tmem1, idesc, first, ready); break;
                default: mma_block<BST, 1, 1, PAIR>(b_full, b_empty, &a_empty[sa], blk, a_lo, b_lo0, b_step, d_tmem, d_tmem1, idesc, first, ready); break;
              }
            }
            if (++sa == p.a_stages) {
              sa = 0;
              pa ^= 1;
            }
          }
        if (elect_one()) umma_commit_x<PAIR>(&tfull[acc]);
      }
    }
  } else if (warp == 2) {
    reg_dec<80>();
  } else {
    // ------------------------------------------------------------ epilogue
    reg_inc<208>();
    const int sub = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    // where this warp's lane 0 reports "accumulators drained": the leader's tempty
    auto release = [&](int acc) {
      if (PAIR && !leader) mbar_arrive_cluster_relaxed(mapa_u32(&tempty[acc], 0));
      else mbar_arrive_relaxed(&tempty[acc]);
    };
    int it = 0;
    for (int item = item0; item < items; item += item_step, ++it) {
      int n_blk, w_blk, h_blk, t;
      decode(item, n_blk, w_blk, h_blk, t);
      const int h = h_blk * CH_T + (m >> 3);
      const int w = w_blk * CH_T + sub * 8 + (m & 7);
      const bool pix_ok = (h < p.H_out) && (w < p.W_out) && (t < p.T_out);
      const int acc = it % p.acc_bufs;
      const uint32_t acc_phase = (it / p.acc_bufs) & 1;
      // halo convs never interleave channels into frames (n_split >= Cout, checked on the host):
      // one 64-bit offset per pixel, no per-chunk divisions
      const long long pix_off = ((static_cast<long long>(t * p.t_mul + p.t_off) * p.H_out + h) * p.W_out + w) *
                                    p.out_C + n_blk * p.NT;
      // Residual (ResidualBlock shortcut, vae:224): the row is pulled towards L2 and its first
      // 32 channels into registers BEFORE waiting for the accumulators, later chunks one chunk
      // ahead — loaded on demand, each chunk stalled the epilogue for a full memory latency
      // (profiles: 21 % of the epilogue warps' time, +29 % kernel time at 96 channels / 720p).
      const bool vec_res = p.residual != nullptr && pix_ok && p.out_mode == 0 && p.vec_ok &&
                           n_blk * p.NT + 32 <= p.Cout;
      uint4 rnext[4] = {};
      if (vec_res) {
        const char* rp = reinterpret_cast<const char*>(p.residual + pix_off);
        for (int bo = 0; bo < p.NT * 2; bo += 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + bo));
        if (p.vec_ok == 2) load_res_chunk32(p.residual + pix_off, rnext);
        else load_res_chunk(p.residual + pix_off, rnext);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             acc * 2 * p.acc_stride + sub * p.acc_stride;
      const int n_base = n_blk * p.NT;
      if (p.norm_out != nullptr) {                     // NT == Cout in {96, 192}, checked on the host
        if (p.NT == 96) epilogue_vec<96, true>(p, t_row, pix_off, 0, pix_ok, vec_res, rnext, release, acc, lane);
        else epilogue_vec<192, true>(p, t_row, pix_off, 0, pix_ok, vec_res, rnext, release, acc, lane);
        continue;                                      // the accumulators were released after pass 1
      } else if (p.out_mode == 0 && p.vec_ok == 2 && n_base + p.NT <= p.Cout &&
                 (p.NT == 96 || p.NT == 128 || p.NT == 192)) {    // warp-uniform
        // full-width tiles of the plain convs (96 / 128 / 192 / 2 x 192 channels): same drain-release-store
        if (p.NT == 96) epilogue_vec<96, false>(p, t_row, pix_off, n_base, pix_ok, vec_res, rnext, release, acc, lane);
        else if (p.NT == 128) {
          // spatial tile index = (t, h_blk, w_blk); only written for tiles that exist
          float* partial = (p.gn_partials != nullptr && t < p.T_out)
                               ? p.gn_partials + ((static_cast<long long>(t) * tiles_h + h_blk) * tiles_w + w_blk) * 64
                               : nullptr;
          if (p.gn_partials != nullptr && partial == nullptr) partial = gsm + 2 * 512;   // odd tail's peer: scratch
          epilogue_vec<128, false>(p, t_row, pix_off, n_base, pix_ok, vec_res, rnext, release, acc, lane,
                                   gsm + (it & 1) * 512, partial, warp - 4);
        }
        else epilogue_vec<192, false>(p, t_row, pix_off, n_base, pix_ok, vec_res, rnext, release, acc, lane);
        continue;
      } else {
        for (int c0 = 0; c0 < p.NT; c0 += 32) {
          const int n0 = n_blk * p.NT + c0;
          if (n0 >= p.Cout) break;                     // warp-uniform
          uint4 rcur[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
          if (vec_res && n0 + 64 <= p.Cout && c0 + 32 < p.NT) load_res_chunk(p.residual + pix_off + c0 + 32, rnext);
          uint32_t rr[32];
          tmem_ld32(t_row + c0, rr);
          tmem_ld_wait();
          if (!pix_ok) continue;
          if (p.vec_ok && n0 + 32 <= p.Cout) {
            uint32_t y[16];
            conv_chunk_packed(rr, p.bias ? p.bias + n0 : nullptr, vec_res ? rcur : nullptr, y);
            store_chunk_packed(reinterpret_cast<bf16*>(p.out) + pix_off + c0, y);
          } else {
            conv_store_chunk_slow(p, rr, t, h, w, n0);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release(acc);
    }
  }
  tc_fence_before();
  if (PAIR) {
    cluster_sync_all();                  // no CTA may exit while its peer can still touch its smem / TMEM
    if (warp == 2) tmem_dealloc_2sm<512>(tmem_base);
  } else {
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
  }
}

bool conv_halo_eligible(int Cin, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                        int T_in, int H_in, int W_in, int T_out, int H_out, int W_out) {
  (void)T_in; (void)T_out; (void)pt;
  return kh == 3 && kw == 3 && (kt == 1 || kt == 3) && st == 1 && sh == 1 && sw == 1 && ph == 1 &&
         pw == 1 && Cin % 16 == 0 && H_out == H_in && W_out == W_in;
}

int conv_halo_launch(const void* x, int T_in, int H_in, int W_in, const void* w_packed, int Cout_pad,
                     ConvParams p, cudaStream_t stream) {
  int NT = 16;
  for (int cand = 256; cand >= 16; cand -= 16)
    if (Cout_pad % cand == 0) { NT = cand; break; }
  p.NT = NT;
  p.n_tiles = Cout_pad / NT;
  p.acc_stride = (NT + 31) & ~31;
  p.acc_bufs = (4 * p.acc_stride <= 512) ? 2 : 1;
  p.desc_mode = 0;
  M4D_REQUIRE(p.out_mode == 1 || p.n_split >= p.Cout, M4D_ERR_UNSUPPORTED);
  if (p.gn_partials != nullptr)
    M4D_REQUIRE(p.norm_out == nullptr && p.vec_ok == 2 && p.out_mode == 0 && NT == 128 && p.Cout == 128 && p.n_tiles == 1,
                M4D_ERR_UNSUPPORTED);
  if (p.norm_out != nullptr) {
    M4D_REQUIRE(p.vec_ok == 2 && (reinterpret_cast<uintptr_t>(p.norm_out) & 31) == 0, M4D_ERR_ALIGN);
    M4D_REQUIRE(p.norm_gamma != nullptr && p.out_mode == 0 && p.n_tiles == 1 && NT == p.Cout &&
                    (NT == 96 || NT == 192) && p.n_split >= p.Cout && p.out_C % 8 == 0,
                M4D_ERR_UNSUPPORTED);
    M4D_REQUIRE(aligned16(p.norm_out) && aligned16(p.norm_gamma) && (p.out == nullptr || aligned16(p.out)) &&
                    (p.residual == nullptr || aligned16(p.residual)),
                M4D_ERR_ALIGN);
  }
  // thin-input mode: Cin in {16, 32} -> 4 / 2 taps per weight box, ring = stages per block
  const int tpb = (p.Cin == 16) ? 4 : (p.Cin == 32 ? 2 : 1);
  const long long n_spatial = static_cast<long long>(p.T_out) * ((p.H_out + CH_T - 1) / CH_T) *
                              ((p.W_out + CH_T - 1) / CH_T);
  // CTA pairs (cta_group::2, one weight stream per two spatial tiles); thin inputs keep the
  // single-CTA kernel
  bool pair = tpb == 1 && NT % 16 == 0 && n_spatial >= 2;
#ifdef M4D_DEV
  if (g_dev_flags & 0x4000) pair = false;
#endif
  const int b_bytes = (pair ? NT / 2 : NT) * 128;       // per CTA and weight stage
  const int fixed = 1024 + 512 + (p.gn_partials ? 5 * 1024 : 0);   // alignment slack + barriers (+ statistics staging)
  // weight ring: all stages of a block when that leaves room for >= 2 halo stages, else 3
  int bst = (CH_SMEM_MAX - fixed - 9 * b_bytes >= 2 * CH_A_STRIDE) ? 9 : 3;
  if (tpb == 2) bst = 5;
  if (tpb == 4) bst = 3;
  p.stages = bst;
  p.a_stages = (CH_SMEM_MAX - fixed - bst * b_bytes) / CH_A_STRIDE;
  if (p.a_stages > CH_MAX_A) p.a_stages = CH_MAX_A;
  M4D_REQUIRE(p.a_stages >= 2, M4D_ERR_UNSUPPORTED);

  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return M4D_ERR_NO_DEVICE;
  if (!aligned16(x) || !aligned16(w_packed)) return M4D_ERR_ALIGN;
  CUtensorMap tmX, tmW;
  {
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(p.Cin), static_cast<cuuint64_t>(W_in),
                          static_cast<cuuint64_t>(H_in), static_cast<cuuint64_t>(T_in)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(p.Cin) * 2, static_cast<cuuint64_t>(W_in) * p.Cin * 2,
                          static_cast<cuuint64_t>(H_in) * W_in * p.Cin * 2};
    cuuint32_t box[4] = {64, CH_HW, CH_HW, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[more4d_b200] conv_halo: cuTensorMapEncodeTiled(X) failed: %d\n", static_cast<int>(r));
      return M4D_ERR_CUDA;
    }
    const long long Ktot = static_cast<long long>(p.kt) * 9 * p.Cin;
    cuuint64_t wdim[2] = {static_cast<cuuint64_t>(Ktot), static_cast<cuuint64_t>(Cout_pad)};
    cuuint64_t wstr[1] = {static_cast<cuuint64_t>(Ktot) * 2};
    cuuint32_t wbox[2] = {64, static_cast<cuuint32_t>(pair ? NT / 2 : NT)};
    cuuint32_t wes[2] = {1, 1};
    r = fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), wdim, wstr, wbox, wes,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[more4d_b200] conv_halo: cuTensorMapEncodeTiled(W) failed: %d\n", static_cast<int>(r));
      return M4D_ERR_CUDA;
    }
  }
  const int smem_bytes = p.a_stages * CH_A_STRIDE + p.stages * b_bytes + fixed;
  void (*kern)(CUtensorMap, CUtensorMap, ConvParams) =
      tpb == 4 ? conv_halo_kernel<3, 4, false> : tpb == 2 ? conv_halo_kernel<5, 2, false>
      : pair ? (bst == 9 ? conv_halo_kernel<9, 1, true> : conv_halo_kernel<3, 1, true>)
      : bst == 9 ? conv_halo_kernel<9, 1, false> : conv_halo_kernel<3, 1, false>;
  {
    int rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_MAX),
                     "cudaFuncSetAttribute(conv_halo)");
    if (rc != M4D_OK) return rc;
  }
  const long long items = (pair ? (n_spatial + 1) / 2 : n_spatial) * p.n_tiles;
  M4D_REQUIRE(items < (1ll << 30), M4D_ERR_BAD_SHAPE);
  const int slots = pair ? sm_count() / 2 : sm_count();          // persistent: one CTA (pair) per SM (pair)
  const int grid = static_cast<int>(items < slots ? items : slots) * (pair ? 2 : 1);
  kern<<<grid, CH_THREADS, smem_bytes, stream>>>(tmX, tmW, p);
  M4D_CHECK_LAUNCH("conv_halo_kernel");
  return M4D_OK;
}

}  // namespace m4d
