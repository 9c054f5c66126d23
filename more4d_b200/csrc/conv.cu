// Causal 3-D / 2-D convolution of the Wan VAE and the trajectory adaptors as an implicit GEMM on
// tcgen05 tensor cores over CHANNELS-LAST activations.
//
// Replaces CausalConv3d (MoRe4D/models/wan_vae.py:21-40: torch.cat with the cache + F.pad +
// cuDNN Conv3d), the Conv2d's of Resample (:82-100) and of the adaptors' ResnetBlocks
// (MoRe4D/models/trajectory_module.py:73-87).  The reference's chunk loop + 2-frame feature
// cache is algebraically a causal convolution over the whole sequence (oracle/vae_oracle.py),
// so the kernel takes the whole [T, H, W, C] sequence and there is no cache, cat, clone or pad
// copy at all: causal (front) padding in time and zero padding in space are TMA
// out-of-bounds zero fill of a 4-D tensor map (negative / overshooting box coordinates).
//
// GEMM view:  M = output pixels (one tile = 8 rows x 16 cols of one output frame = 128 rows),
//             N = output channels (tile NT <= 256),   K = taps x Cin.
// For each tap (a, b, c) and each block of 32 input channels the producer issues ONE TMA box
// {32 ch, 16 w, 8 h, 1 t} at coordinates shifted by the tap — it lands in shared memory as a
// K-major 128 x 32 SWIZZLE_64B operand tile, exactly what tcgen05.mma consumes.  Strided convs
// (downsample) use the tensor map's element strides.  Weights are pre-packed once to
// [Cout_pad, taps * Cin] (tap-major, channel-minor) so the B tile is a plain 2-D TMA box.
//
// CTA = 256 threads, persistent: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (bias, residual add, bf16 rounding like the reference's bf16 path; NHWC
// store, or planar NCTHW store with clamp / sigmoid for the 3-channel model outputs).
#include "conv_common.cuh"

namespace m4d {

constexpr int CV_TH = 8, CV_TW = 16;           // output tile: 8 x 16 pixels = 128 GEMM rows
constexpr int CV_KB = 32;                      // channels per TMA box (64 bytes -> SWIZZLE_64B)
constexpr int CV_A_SUB = 128 * CV_KB * 2;      // 8 KB
constexpr int CV_THREADS = 256;
constexpr int CV_SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t saddr) {
  // K-major SWIZZLE_64B: rows of 64 B, 8-row groups 512 B apart
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

__global__ void __launch_bounds__(CV_THREADS, 1)
conv_cl_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int a_bytes = p.ksub * CV_A_SUB;
  const int b_sub = p.NT * CV_KB * 2;
  const int stage_bytes = a_bytes + p.ksub * b_sub;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_w = (p.W_out + CV_TW - 1) / CV_TW;
  const int tiles_h = (p.H_out + CV_TH - 1) / CV_TH;
  const int tiles = p.T_out * tiles_h * tiles_w * p.n_tiles;
  const int taps = p.kt * p.kh * p.kw;
  const int cblocks = p.Cin / (CV_KB * p.ksub);
  const int ksteps = taps * cblocks;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int r = tile;
        const int n_blk = r % p.n_tiles; r /= p.n_tiles;
        const int w_blk = r % tiles_w; r /= tiles_w;
        const int h_blk = r % tiles_h;
        const int t = r / tiles_h;
        for (int ks = 0; ks < ksteps; ++ks) {
          const int tap = ks / cblocks, cb = ks - tap * cblocks;
          const int a = tap / (p.kh * p.kw), bc = tap - a * (p.kh * p.kw);
          const int b = bc / p.kw, c = bc - b * p.kw;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* a_s = smem + stage * stage_bytes;
          uint8_t* b_s = a_s + a_bytes;
          mbar_arrive_expect_tx(&full[stage], stage_bytes);
          const int x0 = w_blk * CV_TW * p.sw + c - p.pw;
          const int y0 = h_blk * CV_TH * p.sh + b - p.ph;
          const int t0 = t * p.st + a - p.pt;
          for (int s = 0; s < p.ksub; ++s) {
            const int ch = (cb * p.ksub + s) * CV_KB;
            tma_load_4d(a_s + s * CV_A_SUB, &tmX, &full[stage], ch, x0, y0, t0);
            tma_load_2d(b_s + s * b_sub, &tmW, &full[stage], tap * p.Cin + ch, n_blk * p.NT);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.NT, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
          for (int s = 0; s < p.ksub; ++s) {
#pragma unroll
            for (int k = 0; k < CV_KB / 16; ++k) {
              const uint64_t ad = umma_smem_desc_sw64(a_addr + s * CV_A_SUB + k * 32);
              const uint64_t bd = umma_smem_desc_sw64(b_addr + s * b_sub + k * 32);
              umma_ss(d_tmem, ad, bd, idesc, (ks | s | k) != 0);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == p.stages) {
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
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      int r = tile;
      const int n_blk = r % p.n_tiles; r /= p.n_tiles;
      const int w_blk = r % tiles_w; r /= tiles_w;
      const int h_blk = r % tiles_h;
      const int t = r / tiles_h;
      const int h = h_blk * CV_TH + row / CV_TW;
      const int w = w_blk * CV_TW + row % CV_TW;
      const bool pix_ok = (h < p.H_out) && (w < p.W_out);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
      for (int c0 = 0; c0 < p.NT; c0 += 32) {
        const int n0 = n_blk * p.NT + c0;
        if (n0 >= p.Cout) break;                       // warp-uniform
        uint32_t rr[32];
        tmem_ld32(t_row + c0, rr);
        tmem_ld_wait();
        if (!pix_ok) continue;
        conv_store_chunk(p, rr, t, h, w, n0);
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

}  // namespace m4d

using namespace m4d;

static int conv_cl_impl(const void* x, int T_in, int H_in, int W_in, int Cin, const void* w_packed,
                        int Cout, int Cout_pad, const void* bias, int kt, int kh, int kw, int st,
                        int sh, int sw, int pt, int ph, int pw, int T_out, int H_out, int W_out,
                        void* out, int out_C, int t_mul, int t_off, int n_split,
                        const void* residual, int out_mode, int act, const void* skip,
                        const void* norm_gamma, void* norm_out, int norm_silu, float* gn_partials, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && w_packed && (out || norm_out), M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(T_in > 0 && H_in > 0 && W_in > 0 && T_out > 0 && H_out > 0 && W_out > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(Cin > 0 && Cin % 16 == 0, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(Cout > 0 && Cout_pad >= Cout && Cout_pad % 16 == 0, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(kt >= 1 && kh >= 1 && kw >= 1 && st >= 1 && sh >= 1 && sw >= 1 && sh <= 2 && sw <= 2,
              M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(out_mode == 0 || out_mode == 1, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(n_split > 0 && out_C > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(act != 2 || skip != nullptr, M4D_ERR_BAD_SHAPE);

  ConvParams p;
  p.T_out = T_out; p.H_out = H_out; p.W_out = W_out;
  p.Cin = Cin; p.Cout = Cout;
  p.kt = kt; p.kh = kh; p.kw = kw; p.st = st; p.sh = sh; p.sw = sw; p.pt = pt; p.ph = ph; p.pw = pw;
  const int cb32 = Cin / CV_KB;
  p.ksub = (cb32 % 2 == 0) ? 2 : (cb32 % 3 == 0 ? 3 : 1);
  // output-channel tile: largest multiple of 16 <= 256 dividing Cout_pad
  int NT = 16;
  for (int cand = 256; cand >= 16; cand -= 16)
    if (Cout_pad % cand == 0) { NT = cand; break; }
  p.NT = NT;
  p.n_tiles = Cout_pad / NT;
  const int stage_bytes = p.ksub * (CV_A_SUB + NT * CV_KB * 2);
  int stages = CV_SMEM_BUDGET / stage_bytes;
  p.stages = stages > 6 ? 6 : stages;
  M4D_REQUIRE(p.stages >= 2, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(n_split % 32 == 0 || n_split >= Cout, M4D_ERR_UNSUPPORTED);
  p.out = out;
  p.residual = static_cast<const bf16*>(residual);
  p.bias = static_cast<const bf16*>(bias);
  p.out_C = out_C; p.t_mul = t_mul; p.t_off = t_off; p.n_split = n_split;
  p.out_mode = out_mode; p.act = act;
  p.skip = static_cast<const bf16*>(skip);
  p.planar_cstride = static_cast<long long>(T_out) * H_out * W_out;
  p.a_stages = 0; p.acc_bufs = 0; p.acc_stride = 0; p.desc_mode = 0;
  p.norm_gamma = static_cast<const bf16*>(norm_gamma);
  p.norm_out = static_cast<bf16*>(norm_out);
  p.norm_silu = norm_silu;
  p.gn_partials = gn_partials;
  p.vec_ok = (out_mode == 0 && out_C % 8 == 0 && n_split % 8 == 0 && (out == nullptr || aligned16(out)) &&
              (bias == nullptr || aligned16(bias)) && (residual == nullptr || aligned16(residual)))
                 ? 1 : 0;
  // 32-byte aligned pixel rows: the vector epilogues use 256-bit loads / stores
  if (p.vec_ok && out_C % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0 &&
      (reinterpret_cast<uintptr_t>(residual) & 31) == 0 && n_split >= Cout)
    p.vec_ok = 2;

  // 3x3 (x kt) stride-1 convolutions — almost all of the VAE's FLOPs — take the halo-staging
  // kernel (conv_halo.cu); development builds: flag 0x10000 forces the per-tap kernel below.
  bool use_halo = true;
#ifdef M4D_DEV
  use_halo = !(g_dev_flags & 0x10000);
#endif
  if (use_halo &&
      conv_halo_eligible(Cin, kt, kh, kw, st, sh, sw, pt, ph, pw, T_in, H_in, W_in, T_out, H_out, W_out))
    return conv_halo_launch(x, T_in, H_in, W_in, w_packed, Cout_pad, p, stream);
  M4D_REQUIRE(norm_out == nullptr && gn_partials == nullptr, M4D_ERR_UNSUPPORTED);   // fused norms live in conv_halo.cu
  M4D_REQUIRE(Cin % CV_KB == 0, M4D_ERR_UNSUPPORTED);         // per-tap kernel: 32-channel boxes

  CUtensorMap tmX, tmW;
  {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return M4D_ERR_NO_DEVICE;
    if (!aligned16(x)) return M4D_ERR_ALIGN;
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(Cin), static_cast<cuuint64_t>(W_in),
                          static_cast<cuuint64_t>(H_in), static_cast<cuuint64_t>(T_in)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(Cin) * 2, static_cast<cuuint64_t>(W_in) * Cin * 2,
                          static_cast<cuuint64_t>(H_in) * W_in * Cin * 2};
    cuuint32_t box[4] = {CV_KB, static_cast<cuuint32_t>(CV_TW * sw), static_cast<cuuint32_t>(CV_TH * sh), 1};
    cuuint32_t es[4] = {1, static_cast<cuuint32_t>(sw), static_cast<cuuint32_t>(sh), 1};
    CUresult r = fn(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box,
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[more4d_b200] conv: cuTensorMapEncodeTiled(X) failed: %d\n", static_cast<int>(r));
      return M4D_ERR_CUDA;
    }
    if (!aligned16(w_packed)) return M4D_ERR_ALIGN;
    const long long Ktot = static_cast<long long>(kt) * kh * kw * Cin;
    cuuint64_t wdim[2] = {static_cast<cuuint64_t>(Ktot), static_cast<cuuint64_t>(Cout_pad)};
    cuuint64_t wstr[1] = {static_cast<cuuint64_t>(Ktot) * 2};
    cuuint32_t wbox[2] = {CV_KB, static_cast<cuuint32_t>(NT)};
    cuuint32_t wes[2] = {1, 1};
    r = fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), wdim, wstr, wbox,
           wes, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      fprintf(stderr, "[more4d_b200] conv: cuTensorMapEncodeTiled(W) failed: %d\n", static_cast<int>(r));
      return M4D_ERR_CUDA;
    }
  }
  const int smem_bytes = p.stages * stage_bytes + 256 + 1024;
  static int configured = 0;
  if (configured < smem_bytes) {
    int rc = cuda_ok(cudaFuncSetAttribute(conv_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          CV_SMEM_BUDGET + 256 + 1024 + 16 * 1024),
                     "cudaFuncSetAttribute(conv)");
    if (rc != M4D_OK) return rc;
    configured = CV_SMEM_BUDGET + 256 + 1024 + 16 * 1024;
  }
  const long long tiles = static_cast<long long>(T_out) * ((H_out + CV_TH - 1) / CV_TH) *
                          ((W_out + CV_TW - 1) / CV_TW) * p.n_tiles;
  M4D_REQUIRE(tiles < (1ll << 31), M4D_ERR_BAD_SHAPE);
  const int grid = tiles < sm_count() ? static_cast<int>(tiles) : sm_count();
  conv_cl_kernel<<<grid, CV_THREADS, smem_bytes, stream>>>(tmX, tmW, p);
  M4D_CHECK_LAUNCH("conv_cl_kernel");
  return M4D_OK;
}

extern "C" int m4d_conv_cl(const void* x, int T_in, int H_in, int W_in, int Cin, const void* w_packed,
                           int Cout, int Cout_pad, const void* bias, int kt, int kh, int kw, int st,
                           int sh, int sw, int pt, int ph, int pw, int T_out, int H_out, int W_out,
                           void* out, int out_C, int t_mul, int t_off, int n_split,
                           const void* residual, int out_mode, int act, const void* skip,
                           void* stream_) {
  return conv_cl_impl(x, T_in, H_in, W_in, Cin, w_packed, Cout, Cout_pad, bias, kt, kh, kw, st, sh, sw, pt,
                      ph, pw, T_out, H_out, W_out, out, out_C, t_mul, t_off, n_split, residual, out_mode,
                      act, skip, nullptr, nullptr, 0, nullptr, stream_);
}

extern "C" int m4d_conv3x3_rmsnorm_cl(const void* x, int T, int H, int W, int Cin, const void* w_packed,
                                      int Cout, const void* bias, int kt, void* out, const void* residual,
                                      const void* gamma, void* norm_out, int do_silu, void* stream_) {
  M4D_REQUIRE(gamma && norm_out && (kt == 1 || kt == 3), M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(Cout == 96 || Cout == 192, M4D_ERR_UNSUPPORTED);
  return conv_cl_impl(x, T, H, W, Cin, w_packed, Cout, Cout, bias, kt, 3, 3, 1, 1, 1, kt - 1, 1, 1, T, H, W,
                      out, Cout, 1, 0, Cout, residual, 0, 0, nullptr, gamma, norm_out, do_silu, nullptr, stream_);
}

namespace m4d {
// stats[f][slice][64] = sum over the slice's tiles of partials[f * tiles + tile][64], in tile order
// (deterministic: no atomics anywhere on the fused GroupNorm-statistics path)
__global__ void __launch_bounds__(256)
gn_reduce_kernel(const float* __restrict__ partials, float* __restrict__ stats, int tiles) {
  __shared__ float sh[4][64];
  const int f = blockIdx.y, slice = blockIdx.x, slices = gridDim.x;
  const int per = (tiles + slices - 1) / slices;
  const int t0 = slice * per, t1 = min(tiles, t0 + per);
  const int lane64 = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const float* base = partials + static_cast<long long>(f) * tiles * 64 + lane64;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int t = t0 + sub;
  for (; t + 12 < t1; t += 16) {                       // four loads in flight
    a0 += base[static_cast<long long>(t) * 64];
    a1 += base[static_cast<long long>(t + 4) * 64];
    a2 += base[static_cast<long long>(t + 8) * 64];
    a3 += base[static_cast<long long>(t + 12) * 64];
  }
  for (; t < t1; t += 4) a0 += base[static_cast<long long>(t) * 64];
  sh[sub][lane64] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.x < 64)
    stats[(static_cast<long long>(f) * slices + slice) * 64 + threadIdx.x] =
        (sh[0][threadIdx.x] + sh[1][threadIdx.x]) + (sh[2][threadIdx.x] + sh[3][threadIdx.x]);
}
}  // namespace m4d

extern "C" int m4d_conv3x3_gnstats_cl(const void* x, int T, int H, int W, int Cin, const void* w_packed, int Cout,
                                      const void* bias, void* out, const void* residual, float* partials_ws,
                                      float* stats, void* stream_) {
  using namespace m4d;
  M4D_REQUIRE(out && partials_ws && stats, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(Cout == 128, M4D_ERR_UNSUPPORTED);              // 32 groups x 4 channels, one N tile
  int rc = conv_cl_impl(x, T, H, W, Cin, w_packed, Cout, Cout, bias, 1, 3, 3, 1, 1, 1, 0, 1, 1, T, H, W, out, Cout,
                        1, 0, Cout, residual, 0, 0, nullptr, nullptr, nullptr, 0, partials_ws, stream_);
  if (rc != M4D_OK) return rc;
  const int tiles = ((H + 15) / 16) * ((W + 15) / 16);
  M4D_REQUIRE(T <= 65535, M4D_ERR_BAD_SHAPE);
  gn_reduce_kernel<<<dim3(M4D_GN_SLICES, T), 256, 0, static_cast<cudaStream_t>(stream_)>>>(partials_ws, stats, tiles);
  M4D_CHECK_LAUNCH("gn_reduce_kernel");
  return M4D_OK;
}
