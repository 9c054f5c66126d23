// HBM-bound kernels of the VAE / adaptor path on channels-last bf16 activations [P pixels, C]:
// everything between the convolutions.  The reference runs each as several eager passes over
// permuted copies (SURVEY.md K17-K21, §3.3 iv); here each is one read + one write.
// Rounding follows the reference's bf16 inference path (fp32 statistics inside, every tensor
// op's result rounded to bf16 — oracle/vae_oracle.py emulate_bf16).
#include "common.h"
#include "ptx.cuh"

namespace m4d {

__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// RMS_norm (wan_vae.py:43-58: F.normalize over channels * sqrt(C) * gamma) [+ SiLU].
// Every thread owns one 16-byte vector (8 channels) of a pixel, so a pixel is shared by
// LPP = C/8 consecutive threads (12 / 24 / 48 for C = 96 / 192 / 384 — not warp-aligned, hence
// the per-pixel sum of squares is combined through shared memory rather than shuffles).  A
// block covers 256 / LPP pixels per iteration and keeps RN_ITERS independent iterations in
// flight per thread: the first version (one warp per pixel, one load in flight) ran at ~1/5 of
// the HBM roofline (profiles/launches_vae_r01.md).  In place allowed.
constexpr int RN_ITERS = 4;
__global__ void __launch_bounds__(256)
rmsnorm_silu_cl_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, bf16* __restrict__ out,
                       long long pixels, int C, int do_silu) {
  __shared__ float part[RN_ITERS][256];
  __shared__ float inv_s[RN_ITERS][256];
  const int lpp = C >> 3;                         // lanes (16-byte vectors) per pixel
  const int ppb = 256 / lpp;                      // pixels per block-iteration
  const int lp = threadIdx.x / lpp;               // local pixel
  const int lv = threadIdx.x - lp * lpp;          // vector within the pixel
  const bool active = lp < ppb;
  const long long pix0 = (static_cast<long long>(blockIdx.x) * RN_ITERS) * ppb + lp;
  uint4 v[RN_ITERS];
  float ss[RN_ITERS];
#pragma unroll
  for (int it = 0; it < RN_ITERS; ++it) {            // all loads first: RN_ITERS requests in flight per thread
    const long long pix = pix0 + static_cast<long long>(it) * ppb;
    v[it] = (active && pix < pixels) ? __ldcs(reinterpret_cast<const uint4*>(x + pix * C + lv * 8))
                                     : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int it = 0; it < RN_ITERS; ++it) {
    ss[it] = 0.f;
    {
      const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = __uint_as_float(w[e] << 16), b = __uint_as_float(w[e] & 0xFFFF0000u);
        ss[it] += a * a + b * b;
      }
    }
    part[it][threadIdx.x] = ss[it];
  }
  __syncthreads();
  // one thread per pixel folds the pixel's lpp partial sums (every thread doing it cost lpp shared
  // loads per 8 channels: 6 per channel at C = 384) and publishes 1 / norm
  if (active && lv == 0) {
#pragma unroll
    for (int it = 0; it < RN_ITERS; ++it) {
      float tot = 0.f;
      const float* pp = &part[it][lp * lpp];
      for (int i = 0; i < lpp; ++i) tot += pp[i];
      inv_s[it][lp] = 1.0f / fmaxf(bf16_round(sqrtf(tot)), 1e-12f);
    }
  }
  __syncthreads();
  const uint4 g4 = active ? *reinterpret_cast<const uint4*>(gamma + lv * 8) : make_uint4(0, 0, 0, 0);
  const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
  const float sc = sqrtf(static_cast<float>(C));
#pragma unroll
  for (int it = 0; it < RN_ITERS; ++it) {
    const long long pix = pix0 + static_cast<long long>(it) * ppb;
    if (!(active && pix < pixels)) continue;
    const float inv = inv_s[it][lp];
    const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o[e] = rmsnorm_tail_bf16x2(w[e], inv, sc, gw[e]);
      if (do_silu) o[e] = silu_bf16x2(o[e]);
    }
    *reinterpret_cast<uint4*>(out + pix * C + lv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// nearest-exact 2x spatial upsample (wan_vae.py:61-67,82-83) on [T, H, W, C] -> [T, 2H, 2W, C]
__global__ void __launch_bounds__(256)
upsample2x_cl_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int T, int H, int W, int C8) {
  // one thread = one 16-byte vector of one INPUT pixel, read once (streaming) and written to its four
  // output pixels; consecutive threads cover consecutive vectors, so every store instruction of a
  // warp writes whole 512-byte runs of an output row.  (The first version ran one thread per OUTPUT
  // vector: four times the index arithmetic and every input vector fetched four times: 0.60 of the
  // HBM roofline, profiles/rows_r02.md.)
  const long long total = static_cast<long long>(T) * H * W * C8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C8);
  long long r = idx / C8;
  const int wi = static_cast<int>(r % W); r /= W;
  const int hi = static_cast<int>(r % H);
  const long long t = r / H;
  const uint4 v = __ldcs(reinterpret_cast<const uint4*>(x) + idx);
  uint4* o = reinterpret_cast<uint4*>(out) + ((t * 2 * H + 2 * hi) * (2 * W) + 2 * wi) * C8 + c;
  const long long row = static_cast<long long>(2 * W) * C8;
  __stcs(o, v);
  __stcs(o + C8, v);
  __stcs(o + row, v);
  __stcs(o + row + C8, v);
}

// planar [C, T, H, W] -> channels-last [T*H*W, Cpad] (zero-padded channels) with the decoder's
// latent de-normalisation z / inv_std + mean (wan_vae.py:682-686) fused when `div` is given.
__global__ void planar_to_cl_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long P,
                                    int C, int Cpad, const float* __restrict__ div,
                                    const float* __restrict__ add) {
  // one thread = one pixel: C coalesced plane reads, Cpad*2 bytes written as 16-byte vectors
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= P) return;
  uint4* o = reinterpret_cast<uint4*>(out + pix * Cpad);
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      v[e] = 0.f;
      if (c < C) {
        v[e] = __bfloat162float(x[c * P + pix]);
        if (div != nullptr) v[e] = bf16_round(bf16_round(v[e] / div[c]) + add[c]);
      }
    }
    o[c0 >> 3] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                            pack_bf16(v[6], v[7]));
  }
}

// channels-last [P, C] -> planar [C, P] with the encoder's (mu - mean) * inv_std
// (wan_vae.py:540-545) on the first n_affine channels.
__global__ void cl_to_planar_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long P,
                                    int C, int n_affine, const float* __restrict__ sub,
                                    const float* __restrict__ mul) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= P * C) return;
  const long long pix = idx % P;
  const int c = static_cast<int>(idx / P);
  float v = __bfloat162float(x[pix * C + c]);
  if (c < n_affine) v = bf16_round(bf16_round(v - sub[c]) * mul[c]);
  out[idx] = __float2bfloat16_rn(v);
}

// GroupNorm(32 groups, eps 1e-6, affine) + swish for the adaptors (trajectory_module.py:54-60)
// on [F, HW, C] channels-last, statistics per (frame, group).  Pass 1: partial sums.
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats, int HW, int C, int cpg) {
  // grid (chunks, F); thread handles channel-vector v = tid % (C/4) of pixels tid / (C/4) + k*ppb
  __shared__ float sh[2 * 32 * 8];
  const int nvec = C >> 2;
  const int f = blockIdx.y;
  const int v = threadIdx.x % nvec;
  const int ppb = blockDim.x / nvec;
  float s = 0.f, ss = 0.f;
  if (threadIdx.x < ppb * nvec) {
    const int step = gridDim.x * ppb;
    int pix = blockIdx.x * ppb + threadIdx.x / nvec;
    for (; pix + 3 * step < HW; pix += 4 * step) {                  // four loads in flight per thread
      float4 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = ld4(x + (static_cast<long long>(f) * HW + pix + k * step) * C + v * 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s += (a[k].x + a[k].y) + (a[k].z + a[k].w);
        ss += (a[k].x * a[k].x + a[k].y * a[k].y) + (a[k].z * a[k].z + a[k].w * a[k].w);
      }
    }
    for (; pix < HW; pix += step) {
      const float4 a = ld4(x + (static_cast<long long>(f) * HW + pix) * C + v * 4);
      s += (a.x + a.y) + (a.z + a.w);
      ss += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    }
  }
  // a group spans cpg channels = cpg/4 vectors (cpg is a multiple of 4)
  const int g = (v * 4) / cpg;
  for (int i = threadIdx.x; i < 2 * 32 * 8; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  if (threadIdx.x < ppb * nvec) {
    atomicAdd(&sh[(g * 2 + 0) * 8 + (threadIdx.x & 7)], s);
    atomicAdd(&sh[(g * 2 + 1) * 8 + (threadIdx.x & 7)], ss);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[threadIdx.x * 8 + i];
    atomicAdd(&stats[f * 64 + threadIdx.x], t);
  }
}

// Pass 2: y = swish(GroupNorm(x)).  Grid (chunks, F): a block first folds the frame's group
// statistics and the affine into per-channel (A, B) in shared memory — y = x * A + B with
// A = rstd * w, B = b - mean * A — then streams 4 x 256 16-byte vectors per iteration with all
// four loads in flight.  (The first version recomputed mean / rstd per channel pair from global
// memory and paid two 64-bit divisions per vector: 13 ms per 11.6 GB tensor, 3.6 ms is the HBM
// time.)
constexpr int GN_VPT = 8;
__global__ void __launch_bounds__(256)
groupnorm_swish_kernel(const bf16* __restrict__ x, const float* __restrict__ stats, int slices,
                       const bf16* __restrict__ w, const bf16* __restrict__ b, bf16* __restrict__ out,
                       int HW, int C, int cpg, float eps) {
  extern __shared__ float ab[];                   // A[C] | B[C]
  const int f = blockIdx.y;
  const float n = static_cast<float>(HW) * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    float s1 = 0.f, s2 = 0.f;
    for (int sl = 0; sl < slices; ++sl) {              // fixed order: reproducible
      s1 += stats[(f * slices + sl) * 64 + g * 2];
      s2 += stats[(f * slices + sl) * 64 + g * 2 + 1];
    }
    const float mean = s1 / n;
    const float var = fmaxf(s2 / n - mean * mean, 0.f);
    const float a = rsqrtf(var + eps) * __bfloat162float(w[c]);
    ab[c] = a;
    ab[C + c] = __bfloat162float(b[c]) - mean * a;
  }
  __syncthreads();
  const int nvec = C >> 3;
  const int per_frame = HW * nvec;                // 16-byte vectors per frame
  const uint4* xf = reinterpret_cast<const uint4*>(x) + static_cast<long long>(f) * per_frame;
  uint4* of = reinterpret_cast<uint4*>(out) + static_cast<long long>(f) * per_frame;
  // 256 % nvec == 0 (C in {128, 256, ...}): a thread always lands on the same 8 channels, so its
  // (A, B) live in registers and the loop has no shared-memory loads and no index arithmetic
  const bool fixed_c = (256 % nvec) == 0;
  float Ar[8], Br[8];
  if (fixed_c) {
    const int c0 = (threadIdx.x % nvec) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      Ar[e] = ab[c0 + e];
      Br[e] = ab[C + c0 + e];
    }
  }
  for (int i0 = blockIdx.x * (256 * GN_VPT) + threadIdx.x; i0 < per_frame; i0 += gridDim.x * (256 * GN_VPT)) {
    uint4 v[GN_VPT];
#pragma unroll
    for (int k = 0; k < GN_VPT; ++k)
      if (i0 + k * 256 < per_frame) v[k] = __ldcs(xf + i0 + k * 256);
#pragma unroll
    for (int k = 0; k < GN_VPT; ++k) {
      const int i = i0 + k * 256;
      if (i >= per_frame) break;
      const uint32_t xs[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      uint32_t o[4];
      if (fixed_c) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // y = bf16(x * A + B); swish = bf16(y * bf16(sigmoid(y)))   (x * torch.sigmoid(x) on bf16 tensors)
          const uint32_t y = pack_bf16(fmaf(bf16_lo(xs[e]), Ar[2 * e], Br[2 * e]),
                                       fmaf(bf16_hi(xs[e]), Ar[2 * e + 1], Br[2 * e + 1]));
          o[e] = mul_bf16x2(y, sigmoid_bf16x2(y));
        }
      } else {
        const int c0 = (i % nvec) * 8;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 A = *reinterpret_cast<const float2*>(&ab[c0 + 2 * e]);
          const float2 B = *reinterpret_cast<const float2*>(&ab[C + c0 + 2 * e]);
          const uint32_t y = pack_bf16(fmaf(bf16_lo(xs[e]), A.x, B.x), fmaf(bf16_hi(xs[e]), A.y, B.y));
          o[e] = mul_bf16x2(y, sigmoid_bf16x2(y));
        }
      }
      __stcs(of + i, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
}

// row softmax of fp32 logits * scale -> bf16 probabilities (the VAE AttentionBlock's single
// 384-wide head, wan_vae.py:257-261, run as GEMM -> softmax -> GEMM)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int N, long long lds,
                    long long ldp, float scale) {
  __shared__ float red[8];
  const float* sr = s + static_cast<long long>(blockIdx.x) * lds;
  bf16* pr = p + static_cast<long long>(blockIdx.x) * ldp;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, sr[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sum += __expf((sr[i] - mx) * scale);
  sum = wsum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    pr[i] = __float2bfloat16_rn(__expf((sr[i] - mx) * scale) * inv);
}

// Same, the row held in registers: ONE 16-byte-vector read of the logits, one exponential per
// element, 8-byte stores (the three-pass kernel above re-read the row from L1/L2 three times with
// 4-byte loads and evaluated every exponential twice: 0.27 of the HBM roofline at [14400, 14400],
// profiles/rows_r02.md).  N % 4 == 0, N <= 1024 * SR_VPT, rows 16-byte (fp32) / 8-byte (bf16) aligned.
constexpr int SR_VPT = 16;
__global__ void __launch_bounds__(256)
softmax_rows_reg_kernel(const float* __restrict__ s, bf16* __restrict__ p, int N, long long lds,
                        long long ldp, float scale) {
  __shared__ float red[2][8];
  const float4* sr = reinterpret_cast<const float4*>(s + static_cast<long long>(blockIdx.x) * lds);
  uint2* pr = reinterpret_cast<uint2*>(p + static_cast<long long>(blockIdx.x) * ldp);
  const int nv = N >> 2;
  float4 v[SR_VPT];
#pragma unroll
  for (int k = 0; k < SR_VPT; ++k) {
    const int i = threadIdx.x + k * 256;
    v[k] = i < nv ? __ldcs(sr + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < SR_VPT; ++k) mx = fmaxf(fmaxf(mx, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0][0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[0][i]);
  const float c = scale * 1.4426950408889634f;
  const float mc = mx * c;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < SR_VPT; ++k) {
    v[k].x = fast_exp2(fmaf(v[k].x, c, -mc));
    v[k].y = fast_exp2(fmaf(v[k].y, c, -mc));
    v[k].z = fast_exp2(fmaf(v[k].z, c, -mc));
    v[k].w = fast_exp2(fmaf(v[k].w, c, -mc));
    sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  sum = wsum(sum);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[1][i];
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < SR_VPT; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nv) __stcs(pr + i, make_uint2(pack_bf16(v[k].x * inv, v[k].y * inv), pack_bf16(v[k].z * inv, v[k].w * inv)));
  }
}

// out[c, r] = in[r, c]  (bf16), 32x32 smem tiles; in row stride ld_in, out row stride ld_out
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int R, int C,
                                      long long ld_in, long long ld_out) {
  __shared__ bf16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[r * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[c * ld_out + r] = tile[threadIdx.x][i];
  }
}

}  // namespace m4d

using namespace m4d;

extern "C" int m4d_rmsnorm_silu_cl(const void* x, const void* gamma, void* out, long long pixels,
                                   int C, int do_silu, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && gamma && out && pixels > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % 8 == 0 && C <= 2048 && C > 0, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(aligned16(x) && aligned16(out) && aligned16(gamma), M4D_ERR_ALIGN);
  const int ppb = 256 / (C / 8);
  const long long blocks = (pixels + static_cast<long long>(ppb) * RN_ITERS - 1) / (static_cast<long long>(ppb) * RN_ITERS);
  M4D_REQUIRE(blocks < (1ll << 31), M4D_ERR_BAD_SHAPE);
  rmsnorm_silu_cl_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<const bf16*>(gamma), static_cast<bf16*>(out), pixels, C,
      do_silu);
  M4D_CHECK_LAUNCH("rmsnorm_silu_cl_kernel");
  return M4D_OK;
}

extern "C" int m4d_upsample2x_cl(const void* x, void* out, int T, int H, int W, int C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && out && T > 0 && H > 0 && W > 0 && C > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % 8 == 0 && aligned16(x) && aligned16(out), M4D_ERR_ALIGN);
  const long long total = static_cast<long long>(T) * H * W * (C / 8);
  const long long blocks = (total + 255) / 256;
  M4D_REQUIRE(blocks < (1ll << 31), M4D_ERR_BAD_SHAPE);
  upsample2x_cl_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), T, H, W, C / 8);
  M4D_CHECK_LAUNCH("upsample2x_cl_kernel");
  return M4D_OK;
}

extern "C" int m4d_planar_to_cl(const void* x, void* out, long long P, int C, int Cpad,
                                const float* div, const float* add, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && out && P > 0 && C > 0 && Cpad >= C, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE((div == nullptr) == (add == nullptr), M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(Cpad % 8 == 0 && aligned16(out), M4D_ERR_ALIGN);
  const long long blocks = (P + 255) / 256;
  M4D_REQUIRE(blocks < (1ll << 31), M4D_ERR_BAD_SHAPE);
  planar_to_cl_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), P, C, Cpad, div, add);
  M4D_CHECK_LAUNCH("planar_to_cl_kernel");
  return M4D_OK;
}

extern "C" int m4d_cl_to_planar(const void* x, void* out, long long P, int C, int n_affine,
                                const float* sub, const float* mul, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && out && P > 0 && C > 0 && n_affine >= 0 && n_affine <= C, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(n_affine == 0 || (sub && mul), M4D_ERR_BAD_SHAPE);
  const long long blocks = (P * C + 255) / 256;
  M4D_REQUIRE(blocks < (1ll << 31), M4D_ERR_BAD_SHAPE);
  cl_to_planar_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), P, C, n_affine, sub, mul);
  M4D_CHECK_LAUNCH("cl_to_planar_kernel");
  return M4D_OK;
}

extern "C" int m4d_groupnorm_swish_cl(const void* x, const void* weight, const void* bias, void* out,
                                      float* stats_ws, int F, int HW, int C, int groups, float eps,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && weight && bias && out && stats_ws && F > 0 && HW > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(groups == 32 && C % (groups * 4) == 0 && C <= 1024, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(aligned16(x) && aligned16(out) && aligned16(weight) && aligned16(bias), M4D_ERR_ALIGN);
  M4D_REQUIRE(F <= 65535, M4D_ERR_BAD_SHAPE);
  int rc = cuda_ok(cudaMemsetAsync(stats_ws, 0, sizeof(float) * 64 * F, stream), "memset(groupnorm stats)");
  if (rc != M4D_OK) return rc;
  const int cpg = C / groups;
  const int ppb = 256 / (C / 4);
  int chunks = (HW + ppb * 16 - 1) / (ppb * 16);
  if (chunks > 1024) chunks = 1024;
  if (chunks < 1) chunks = 1;
  groupnorm_stats_kernel<<<dim3(chunks, F), 256, 0, stream>>>(static_cast<const bf16*>(x), stats_ws, HW, C, cpg);
  M4D_CHECK_LAUNCH("groupnorm_stats_kernel");
  M4D_REQUIRE(static_cast<long long>(HW) * (C / 8) < (1ll << 31), M4D_ERR_BAD_SHAPE);
  const int per_frame = HW * (C / 8);
  int gx = (per_frame + 256 * GN_VPT - 1) / (256 * GN_VPT);
  const int cap = (8 * sm_count() + F - 1) / F > 1 ? (8 * sm_count() + F - 1) / F : 1;   // ~8 blocks per SM
  if (gx > cap) gx = cap;
  groupnorm_swish_kernel<<<dim3(gx, F), 256, 2 * C * sizeof(float), stream>>>(
      static_cast<const bf16*>(x), stats_ws, 1, static_cast<const bf16*>(weight), static_cast<const bf16*>(bias),
      static_cast<bf16*>(out), HW, C, cpg, eps);
  M4D_CHECK_LAUNCH("groupnorm_swish_kernel");
  return M4D_OK;
}

extern "C" int m4d_groupnorm_apply_cl(const void* x, const void* weight, const void* bias, void* out,
                                      const float* stats, int slices, int F, int HW, int C, int groups, float eps,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && weight && bias && out && stats && F > 0 && HW > 0 && slices > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(groups == 32 && C % (groups * 4) == 0 && C <= 1024, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(aligned16(x) && aligned16(out) && aligned16(weight) && aligned16(bias), M4D_ERR_ALIGN);
  M4D_REQUIRE(F <= 65535 && static_cast<long long>(HW) * (C / 8) < (1ll << 31), M4D_ERR_BAD_SHAPE);
  const int per_frame = HW * (C / 8);
  int gx = (per_frame + 256 * GN_VPT - 1) / (256 * GN_VPT);
  const int cap = (8 * sm_count() + F - 1) / F > 1 ? (8 * sm_count() + F - 1) / F : 1;
  if (gx > cap) gx = cap;
  groupnorm_swish_kernel<<<dim3(gx, F), 256, 2 * C * sizeof(float), stream>>>(
      static_cast<const bf16*>(x), stats, slices, static_cast<const bf16*>(weight), static_cast<const bf16*>(bias),
      static_cast<bf16*>(out), HW, C, C / groups, eps);
  M4D_CHECK_LAUNCH("groupnorm_swish_kernel");
  return M4D_OK;
}

extern "C" int m4d_softmax_rows(const float* s, void* p, int rows, int N, long long lds, long long ldp,
                                float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(s && p && rows > 0 && N > 0 && lds >= N && ldp >= N, M4D_ERR_BAD_SHAPE);
  if (N % 4 == 0 && N <= 1024 * SR_VPT && lds % 4 == 0 && ldp % 4 == 0 && aligned16(s) &&
      (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
    softmax_rows_reg_kernel<<<rows, 256, 0, stream>>>(s, static_cast<bf16*>(p), N, lds, ldp, scale);
    M4D_CHECK_LAUNCH("softmax_rows_reg_kernel");
    return M4D_OK;
  }
  softmax_rows_kernel<<<rows, 256, 0, stream>>>(s, static_cast<bf16*>(p), N, lds, ldp, scale);
  M4D_CHECK_LAUNCH("softmax_rows_kernel");
  return M4D_OK;
}

extern "C" int m4d_transpose_bf16(const void* in, void* out, int R, int C, long long ld_in,
                                  long long ld_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(in && out && R > 0 && C > 0 && ld_in >= C && ld_out >= R, M4D_ERR_BAD_SHAPE);
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  M4D_REQUIRE(grid.y <= 65535, M4D_ERR_BAD_SHAPE);
  transpose_bf16_kernel<<<grid, dim3(32, 8), 0, stream>>>(static_cast<const bf16*>(in), static_cast<bf16*>(out),
                                                          R, C, ld_in, ld_out);
  M4D_CHECK_LAUNCH("transpose_bf16_kernel");
  return M4D_OK;
}
