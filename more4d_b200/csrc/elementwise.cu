// HBM-bound row kernels of the DiT block: everything between the GEMMs and the attention.
// Each fuses what the reference runs as 3-10 eager kernels into one pass:
// one read and one write of the activation, 128-bit vectorised, one CTA per token row.
#include "common.h"
#include "ptx.cuh"

namespace m4d {

constexpr int ROW_THREADS = 256;
constexpr int ROW_MAXV = 8;    // float4 / 4xbf16 vectors per thread -> C <= 8192

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA; every thread gets the result.  `red` is 8 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < ROW_THREADS / 32) ? red[lane] : 0.f;
  return warp_sum(t);
}

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ void store4(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16(v.x, v.y);
  u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------------------------------------------------------------------------------
// LayerNorm (fp32 statistics, two-pass) [* weight + bias] [* (1 + scale) + shift] -> bf16/fp32
//   WanLayerNorm t4d:397-407 + AdaLN modulation t4d:662,677,720; norm3 (affine) t4d:674;
//   MLPProj's two LayerNorms t4d:729-733.
//   Optional Motion-Perception-Module injection (SpatialGuidanceModule t4d:757-783):
//   y = y * (1 + bf16(sg_scale*gate)) + bf16(sg_shift*gate) for rows < sg_rows of each batch.
// ---------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_modulate_kernel(const TIn* __restrict__ x, const bf16* __restrict__ weight,
                          const bf16* __restrict__ bias, const float* __restrict__ shift,
                          const float* __restrict__ scale, long long mod_bstride,
                          int rows_per_batch, int C, float eps, TOut* __restrict__ out,
                          const bf16* __restrict__ sg, long long sg_bstride, int sg_rows,
                          const bf16* __restrict__ sg_gate) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const TIn* xr = x + row * C;
  const int nvec = C >> 2;
  float4 v[ROW_MAXV];
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {          // every load of the row in flight before the first use
    const int vi = threadIdx.x + i * ROW_THREADS;
    v[i] = vi < nvec ? load4(xr + vi * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = block_sum(s, red) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
    const int vi = threadIdx.x + i * ROW_THREADS;
    if (vi < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / C + eps);
  const long long batch = row / rows_per_batch;
  const int row_in_batch = static_cast<int>(row - batch * rows_per_batch);
  const float* sh = shift ? shift + batch * mod_bstride : nullptr;
  const float* sc = scale ? scale + batch * mod_bstride : nullptr;
  const bf16* sgr = (sg != nullptr && row_in_batch < sg_rows)
                        ? sg + batch * sg_bstride + static_cast<long long>(row_in_batch) * 2 * C
                        : nullptr;
  TOut* orow = out + row * C;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
    const int vi = threadIdx.x + i * ROW_THREADS;
    if (vi < nvec) {
      const int c0 = vi * 4;
      float4 y = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd,
                             (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
      if (weight) {
        const float4 w = load4(weight + c0);
        y.x *= w.x; y.y *= w.y; y.z *= w.z; y.w *= w.w;
      }
      if (bias) {
        const float4 b = load4(bias + c0);
        y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
      }
      if (sc) {
        const float4 a = load4(sc + c0);
        y.x *= 1.f + a.x; y.y *= 1.f + a.y; y.z *= 1.f + a.z; y.w *= 1.f + a.w;
      }
      if (sh) {
        const float4 a = load4(sh + c0);
        y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w;
      }
      if (sgr) {
        const float4 gs = load4(sgr + c0), gh = load4(sgr + C + c0), g = load4(sg_gate + c0);
        y.x = y.x * bf16_round(1.f + bf16_round(gs.x * g.x)) + bf16_round(gh.x * g.x);
        y.y = y.y * bf16_round(1.f + bf16_round(gs.y * g.y)) + bf16_round(gh.y * g.y);
        y.z = y.z * bf16_round(1.f + bf16_round(gs.z * g.z)) + bf16_round(gh.z * g.z);
        y.w = y.w * bf16_round(1.f + bf16_round(gs.w * g.w)) + bf16_round(gh.w * g.w);
      }
      store4(orow + c0, y);
    }
  }
}

// The DiT's hot instance — fp32 residual stream in, bf16 out, C <= 5120, no Motion-Perception
// injection — with FOUR WARPS PER ROW, two rows per CTA: a lane keeps 10 float4 of its row in
// registers, the two statistics are reduced with shuffles + one shared-memory exchange under a
// 128-thread named barrier each (the CTA-per-row kernel above pays four __syncthreads of 256
// threads per row and a CTA launch per 20 KB: 4.5 TB/s = 0.70 of the HBM copy peak,
// profiles/rows_r02.md).  Same arithmetic, same summation tree per lane.
constexpr int LW_MAXV = 10;
__global__ void __launch_bounds__(256, 3)
layernorm_modulate_wg_kernel(const float* __restrict__ x, const bf16* __restrict__ weight,
                             const bf16* __restrict__ bias, const float* __restrict__ shift,
                             const float* __restrict__ scale, long long mod_bstride, long long rows,
                             int rows_per_batch, int C, float eps, bf16* __restrict__ out) {
  __shared__ float red[2][2][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp >> 2, wq = warp & 3;
  long long row = static_cast<long long>(blockIdx.x) * 2 + slot;
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const float* xr = x + row * C;
  const int nvec = C >> 2;
  float4 v[LW_MAXV];
#pragma unroll
  for (int i = 0; i < LW_MAXV; ++i) {
    const int vi = lane + (4 * i + wq) * 32;
    v[i] = vi < nvec ? __ldcs(reinterpret_cast<const float4*>(xr + vi * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LW_MAXV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  s = warp_sum(s);
  if (lane == 0) red[slot][0][wq] = s;
  named_bar_sync(1 + slot, 128);
  const float mean = ((red[slot][0][0] + red[slot][0][1]) + (red[slot][0][2] + red[slot][0][3])) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LW_MAXV; ++i) {
    const int vi = lane + (4 * i + wq) * 32;
    if (vi < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) red[slot][1][wq] = ss;
  named_bar_sync(1 + slot, 128);
  const float rstd = rsqrtf(((red[slot][1][0] + red[slot][1][1]) + (red[slot][1][2] + red[slot][1][3])) / C + eps);
  const long long batch = row / rows_per_batch;
  const float* sh = shift ? shift + batch * mod_bstride : nullptr;
  const float* sc = scale ? scale + batch * mod_bstride : nullptr;
  bf16* orow = out + row * C;
#pragma unroll
  for (int i = 0; i < LW_MAXV; ++i) {
    const int vi = lane + (4 * i + wq) * 32;
    if (live && vi < nvec) {
      const int c0 = vi * 4;
      float4 y = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd,
                             (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
      if (weight) {
        const float4 w = load4(weight + c0);
        y.x *= w.x; y.y *= w.y; y.z *= w.z; y.w *= w.w;
      }
      if (bias) {
        const float4 b = load4(bias + c0);
        y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
      }
      if (sc) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(sc + c0));
        y.x *= 1.f + a.x; y.y *= 1.f + a.y; y.z *= 1.f + a.z; y.w *= 1.f + a.w;
      }
      if (sh) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(sh + c0));
        y.x += a.x; y.y += a.y; y.z += a.z; y.w += a.w;
      }
      store4(orow + c0, y);
    }
  }
}

// ---------------------------------------------------------------------------------------
// WanRMSNorm over the full channel dim + 3-axis RoPE, in place on bf16 [B, L, heads*128].
//   t4d:378-394: y = bf16(bf16(x * bf16(rstd)) * w), rstd = rsqrt(mean(x^2) + eps) in fp32;
//   t4d:340-369: rotate adjacent pairs by the token's (frame | row | col) angle; tokens at
//   positions >= F*H*W pass through.  cos/sin tables are the fp32 roundings of the reference's
//   float64 `freqs` table [1024, 64].
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS)
rmsnorm_rope_kernel(bf16* __restrict__ x, const bf16* __restrict__ weight,
                    const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                    const int* __restrict__ grid_fhw, int L, int C, int head_dim, float eps,
                    long long row_stride) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const int b = static_cast<int>(row / L);
  const int l = static_cast<int>(row - static_cast<long long>(b) * L);
  bf16* xr = x + row * row_stride;
  const int nvec = C >> 2;
  float4 v[ROW_MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
    const int vi = threadIdx.x + i * ROW_THREADS;
    if (vi < nvec) {
      v[i] = load4(xr + vi * 4);
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  float rstd = 1.f;
  if (weight != nullptr) rstd = bf16_round(rsqrtf(block_sum(ss, red) / C + eps));

  bool rope = false;
  int pf = 0, ph = 0, pw = 0;
  if (rope_cos != nullptr) {
    const int F = grid_fhw[b * 3 + 0], H = grid_fhw[b * 3 + 1], W = grid_fhw[b * 3 + 2];
    if (l < F * H * W) {
      rope = true;
      pf = l / (H * W);
      const int rem = l - pf * H * W;
      ph = rem / W;
      pw = rem - ph * W;
    }
  }
  const int half = head_dim >> 1;               // pairs per head (64)
  const int n_hw = head_dim / 6;                // 21 pairs each for row / col
  const int n_f = half - 2 * n_hw;              // 22 pairs for the frame axis
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
    const int vi = threadIdx.x + i * ROW_THREADS;
    if (vi < nvec) {
      const int c0 = vi * 4;
      float4 y = v[i];
      if (weight != nullptr) {
        const float4 w = load4(weight + c0);
        y.x = bf16_round(bf16_round(y.x * rstd) * w.x);
        y.y = bf16_round(bf16_round(y.y * rstd) * w.y);
        y.z = bf16_round(bf16_round(y.z * rstd) * w.z);
        y.w = bf16_round(bf16_round(y.w * rstd) * w.w);
      }
      if (rope) {
        const int pair0 = (c0 % head_dim) >> 1;   // two pairs: pair0, pair0 + 1
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int pi = pair0 + k;
          const int pos = pi < n_f ? pf : (pi < n_f + n_hw ? ph : pw);
          const float cs = __ldg(rope_cos + pos * half + pi);
          const float sn = __ldg(rope_sin + pos * half + pi);
          float& re = k == 0 ? y.x : y.z;
          float& im = k == 0 ? y.y : y.w;
          const float r2 = re * cs - im * sn;
          const float i2 = re * sn + im * cs;
          re = r2;
          im = i2;
        }
      }
      store4(xr + c0, y);
    }
  }
}

// Same arithmetic, TWO WARPS PER ROW with 16-byte vectors — the shape the DiT runs (head_dim 128,
// C <= 5120).  The CTA-per-row kernel above moves 8 bytes per load, five loads per thread, and
// pays two block reductions per row: 2.6 TB/s = 0.40 of the measured HBM copy peak at
// [2, 50400, 5120] (profiles/rows_r02.md).  Here a lane owns every other 32-vector group of its row (up to 10 loads of 16 B in flight per
// lane, ~90 registers, three CTAs per SM), the sum of squares is reduced with shuffles plus one
// shared-memory exchange between the two warps, and — because 32 vectors span exactly two 128-channel heads — every vector of a
// lane covers the SAME four rotary pairs (lane % 16) * 4 .. + 3 of some head: the lane fetches its
// four (cos, sin) pairs once per row instead of two table loads per pair and head.
constexpr int RW_MAXV = 10;          // 16-byte vectors per lane: C <= 2 warps * 32 lanes * 8 * 10 = 5120
constexpr int RW_WARPS = 8;          // per CTA: 4 rows x 2 warps
template <bool NORM>
__global__ void __launch_bounds__(RW_WARPS * 32, 2)
rmsnorm_rope_warp_kernel(bf16* __restrict__ x, const bf16* __restrict__ weight,
                         const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                         const int* __restrict__ grid_fhw, long long rows, int L, int C, float eps,
                         long long row_stride) {
  __shared__ float part[RW_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wr = warp & 1;                       // which half of the row's vectors this warp owns
  long long row = static_cast<long long>(blockIdx.x) * (RW_WARPS / 2) + (warp >> 1);
  const bool live = row < rows;
  if (!live) row = rows - 1;                     // keep the warp in step with the CTA barrier below
  const int b = static_cast<int>(row / L);
  const int l = static_cast<int>(row - static_cast<long long>(b) * L);
  bf16* xr = x + row * row_stride;
  const int nvec = C >> 3;
  uint4 v[RW_MAXV];
  // all loads of the row are issued before the first use: with load and accumulate in one loop the
  // compiler kept them in program order and a row cost ~20 dependent DRAM round trips (measured:
  // 2.4 TB/s, profiles/rows_r02.md, first version of this kernel)
#pragma unroll
  for (int i = 0; i < RW_MAXV; ++i) {
    const int vi = lane + (2 * i + wr) * 32;
    v[i] = vi < nvec ? __ldcs(reinterpret_cast<const uint4*>(xr + vi * 8)) : make_uint4(0, 0, 0, 0);
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < RW_MAXV; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = __uint_as_float(w[e] << 16), hi = __uint_as_float(w[e] & 0xFFFF0000u);
      ss += lo * lo + hi * hi;
    }
  }
  float rstd = 1.f;
  if (NORM) {
    ss = warp_sum(ss);
    if (lane == 0) part[warp] = ss;
    named_bar_sync(1 + (warp >> 1), 64);           // the row's two warps only; other rows do not wait
    rstd = bf16_round(rsqrtf((part[warp & ~1] + part[warp | 1]) / C + eps));   // same order in both warps
  }

  bool rope = false;
  float cs[4], sn[4];
  if (rope_cos != nullptr) {
    const int F = grid_fhw[b * 3 + 0], H = grid_fhw[b * 3 + 1], W = grid_fhw[b * 3 + 2];
    if (l < F * H * W) {
      rope = true;
      const int pf = l / (H * W);
      const int rem = l - pf * H * W;
      const int ph = rem / W;
      const int pw = rem - ph * W;
      const int pair0 = (lane & 15) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int pi = pair0 + k;                  // 22 frame | 21 row | 21 col pairs (t4d:346,928-935)
        const int pos = pi < 22 ? pf : (pi < 43 ? ph : pw);
        cs[k] = __ldg(rope_cos + pos * 64 + pi);
        sn[k] = __ldg(rope_sin + pos * 64 + pi);
      }
    }
  }
  const uint32_t rstd2 = pack_bf16(rstd, rstd);
  // the norm weight (10 KB) stays L1 resident; 16 warps per SM cover its latency
#pragma unroll
  for (int i = 0; i < RW_MAXV; ++i) {
    const int vi = lane + (2 * i + wr) * 32;
    if (live && vi < nvec) {
      uint4 gw = make_uint4(0, 0, 0, 0);
      if (NORM) gw = __ldg(reinterpret_cast<const uint4*>(weight + vi * 8));
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t g[4] = {gw.x, gw.y, gw.z, gw.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // bf16(bf16(x * rstd) * w) with rstd already rounded to bf16 (t4d:378-394): two packed multiplies
        const uint32_t xn = NORM ? mul_bf16x2(mul_bf16x2(w[e], rstd2), g[e]) : w[e];
        float re = bf16_lo(xn), im = bf16_hi(xn);
        if (rope) {
          const float r2 = re * cs[e] - im * sn[e];
          const float i2 = re * sn[e] + im * cs[e];
          re = r2;
          im = i2;
        }
        o[e] = pack_bf16(re, im);
      }
      *reinterpret_cast<uint4*>(xr + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// y[M,N] (fp32) = act(x[M,K] fp32) . W[N,K]^T (bf16) + b   for M <= 8 rows: the time-embedding
// MLPs, which the reference runs under autocast(float32) (t4d:1160-1171).  HBM-bound on W:
// one warp per output column, 128-bit weight loads, fp32 accumulation.
// ---------------------------------------------------------------------------------------
template <int ACT_IN>   // 0 none, 1 SiLU on x
__global__ void __launch_bounds__(256)
small_linear_f32_kernel(const float* __restrict__ x, const bf16* __restrict__ w,
                        const bf16* __restrict__ bias, float* __restrict__ y, int M, int N, int K,
                        int act_out) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  const bf16* wr = w + static_cast<long long>(n) * K;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + k0));
    const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
    float wf[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      wf[2 * e] = __uint_as_float(ww[e] << 16);
      wf[2 * e + 1] = __uint_as_float(ww[e] & 0xFFFF0000u);
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < M) {
        const float4 a = load4(x + static_cast<long long>(m) * K + k0);
        const float4 c = load4(x + static_cast<long long>(m) * K + k0 + 4);
        float xs[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float xv = xs[e];
          if (ACT_IN == 1) xv = xv / (1.f + __expf(-xv));
          acc[m] = fmaf(xv, wf[e], acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0) {
    const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
    for (int m = 0; m < M; ++m) {
      float r = acc[m] + bv;
      if (act_out == 1) r = r / (1.f + __expf(-r));
      y[static_cast<long long>(m) * N + n] = r;
    }
  }
}

// sinusoidal timestep embedding in float64 (t4d:239-249), stored as fp32 [B, dim]: cos | sin
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim,
                                          float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx - b * half;
  const double freq = pow(10000.0, -static_cast<double>(i) / half);
  const double a = static_cast<double>(t[b]) * freq;
  out[b * dim + i] = static_cast<float>(cos(a));
  out[b * dim + half + i] = static_cast<float>(sin(a));
}

// out[b, i] = a[i] + e[b, i % m]   (modulation + e0 with m == n, t4d:659; head modulation +
// e.unsqueeze(1) with n == 2m, t4d:717)
__global__ void add_bcast_f32_kernel(const bf16* __restrict__ a, const float* __restrict__ e,
                                     float* __restrict__ out, int B, long long n, long long m) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * n) return;
  const long long b = idx / n, i = idx - b * n;
  out[idx] = __bfloat162float(a[i]) + e[b * m + i % m];
}

// ---------------------------------------------------------------------------------------
// patchify: channel-concat of x [B,Cx,T,H,W] and y [B,Cy,T,H,W] (t4d:1069-1070) gathered into
// the im2col matrix of the (1,2,2)/stride-(1,2,2) patch conv (t4d:1073): rows = tokens in
// (frame,row,col) order, cols = (channel, kh, kw) — the order of the conv weight [C,Cin,1,2,2].
// ---------------------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, int Cx,
                                int Cy, int T, int H, int W, bf16* __restrict__ out) {
  // one thread per (token, channel): writes 4 contiguous bf16
  const int Hp = H >> 1, Wp = W >> 1;
  const int Cin = Cx + Cy;
  const long long total = static_cast<long long>(gridDim.y) * T * Hp * Wp * Cin;
  long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const long long per_b = static_cast<long long>(T) * Hp * Wp * Cin;
  if (idx >= per_b) return;
  (void)total;
  const int wq = static_cast<int>(idx % Wp);
  long long r = idx / Wp;
  const int c = static_cast<int>(r % Cin);
  r /= Cin;
  const int hq = static_cast<int>(r % Hp);
  const int t = static_cast<int>(r / Hp);
  const bf16* src = c < Cx ? x + ((static_cast<long long>(b) * Cx + c) * T + t) * H * W
                           : y + ((static_cast<long long>(b) * Cy + (c - Cx)) * T + t) * H * W;
  const bf16* s0 = src + static_cast<long long>(2 * hq) * W + 2 * wq;
  const uint32_t top = *reinterpret_cast<const uint32_t*>(s0);
  const uint32_t bot = *reinterpret_cast<const uint32_t*>(s0 + W);
  const long long tok = (static_cast<long long>(t) * Hp + hq) * Wp + wq;
  bf16* dst = out + (static_cast<long long>(b) * T * Hp * Wp + tok) * (Cin * 4) + c * 4;
  *reinterpret_cast<uint2*>(dst) = make_uint2(top, bot);
}

// unpatchify (t4d:1343-1366): tokens [B, L, 4*Cout] (cols = (kh,kw,c)) -> [B, Cout, T, H, W]
__global__ void unpatchify_kernel(const bf16* __restrict__ tok, long long tok_bstride,
                                  int skip_tokens, int Cout, int T, int H, int W,
                                  bf16* __restrict__ out) {
  const int Hp = H >> 1, Wp = W >> 1;
  const long long per_b = static_cast<long long>(Cout) * T * H * W;
  long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= per_b) return;
  const int w = static_cast<int>(idx % W);
  long long r = idx / W;
  const int h = static_cast<int>(r % H);
  r /= H;
  const int t = static_cast<int>(r % T);
  const int c = static_cast<int>(r / T);
  const long long token = skip_tokens + (static_cast<long long>(t) * Hp + (h >> 1)) * Wp + (w >> 1);
  const int col = ((h & 1) * 2 + (w & 1)) * Cout + c;
  out[b * per_b + idx] = tok[b * tok_bstride + token * (4 * Cout) + col];
}

// fp32 <- bf16 widening copy with row remap: dst[b, dst_row0 + r, :] = src[b*rows + r, :]
__global__ void widen_rows_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int rows,
                                  int C, long long dst_bstride, int dst_row0) {
  const long long n4 = static_cast<long long>(rows) * (C >> 2);
  long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= n4) return;
  const long long r = idx / (C >> 2);
  const int c0 = static_cast<int>(idx - r * (C >> 2)) * 4;
  const float4 v = load4(src + (static_cast<long long>(b) * rows + r) * C + c0);
  store4(dst + b * dst_bstride + (dst_row0 + r) * C + c0, v);
}

// Classifier-free guidance + flow-matching Euler step (pctl:820-825) on bf16 latents, with the
// reference's rounding: the CFG combine is bf16 arithmetic, the scheduler step is fp32 on the
// upcast sample and the result is cast back to bf16.
__global__ void cfg_euler_step_kernel(const bf16* __restrict__ uncond, const bf16* __restrict__ text,
                                      bf16* __restrict__ latents, float guidance, float dt,
                                      long long n) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const float u = __bfloat162float(uncond[idx]);
  const float tx = __bfloat162float(text[idx]);
  const float d = bf16_round(tx - u);
  const float np = bf16_round(u + bf16_round(guidance * d));
  const float x = __bfloat162float(latents[idx]);
  latents[idx] = __float2bfloat16_rn(x + dt * np);
}

// out bf16 = SiLU(x fp32): prologue of the Motion-Perception-Module projection (t4d:746-748)
__global__ void silu_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, long long n4) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  float4 v = load4(x + idx * 4);
  v.x = v.x / (1.f + __expf(-v.x));
  v.y = v.y / (1.f + __expf(-v.y));
  v.z = v.z / (1.f + __expf(-v.z));
  v.w = v.w / (1.f + __expf(-v.w));
  store4(out + idx * 4, v);
}

#ifdef M4D_DEV
int g_dev_flags = 0;
#endif

}  // namespace m4d

using namespace m4d;

extern "C" int m4d_layernorm_modulate(const void* x, int x_is_bf16, const void* weight,
                                      const void* bias, const float* shift, const float* scale,
                                      long long mod_batch_stride, long long rows, int rows_per_batch,
                                      int C, float eps, void* out, int out_is_f32,
                                      const void* sg, long long sg_batch_stride, int sg_rows,
                                      const void* sg_gate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && out && rows > 0 && C > 0 && rows_per_batch > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % 4 == 0 && C <= ROW_THREADS * 4 * ROW_MAXV, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(aligned16(x) && aligned16(out) && mod_batch_stride % 4 == 0, M4D_ERR_ALIGN);
  M4D_REQUIRE(rows < (1ll << 31), M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(sg == nullptr || sg_gate != nullptr, M4D_ERR_BAD_SHAPE);
  const bf16* w = static_cast<const bf16*>(weight);
  const bf16* bb = static_cast<const bf16*>(bias);
  const bf16* sgp = static_cast<const bf16*>(sg);
  const bf16* sgg = static_cast<const bf16*>(sg_gate);
  if (!x_is_bf16 && !out_is_f32 && sgp == nullptr && C <= 128 * 4 * LW_MAXV &&
      (shift == nullptr || aligned16(shift)) && (scale == nullptr || aligned16(scale))) {
    layernorm_modulate_wg_kernel<<<static_cast<unsigned>((rows + 1) / 2), 256, 0, stream>>>(
        static_cast<const float*>(x), w, bb, shift, scale, mod_batch_stride, rows, rows_per_batch, C, eps,
        static_cast<bf16*>(out));
    M4D_CHECK_LAUNCH("layernorm_modulate_wg_kernel");
    return M4D_OK;
  }
  dim3 grid(static_cast<unsigned>(rows));
#define LAUNCH_LN(TI, TO)                                                                       \
  layernorm_modulate_kernel<TI, TO><<<grid, ROW_THREADS, 0, stream>>>(                          \
      static_cast<const TI*>(x), w, bb, shift, scale, mod_batch_stride, rows_per_batch, C, eps, \
      static_cast<TO*>(out), sgp, sg_batch_stride, sg_rows, sgg)
  if (x_is_bf16 && out_is_f32) LAUNCH_LN(bf16, float);
  else if (x_is_bf16) LAUNCH_LN(bf16, bf16);
  else if (out_is_f32) LAUNCH_LN(float, float);
  else LAUNCH_LN(float, bf16);
#undef LAUNCH_LN
  M4D_CHECK_LAUNCH("layernorm_modulate_kernel");
  return M4D_OK;
}

namespace m4d {

// WanRMSNorm over the full channel dim (wan_transformer4d.py:378-394) whose output is SCATTERED
// by head group into P destination buffers — the sending half of the Ulysses all-to-all of the
// sequence-parallel self-attention, fused into the normalisation pass: destination g receives
// channels [g*C/P, (g+1)*C/P) of every local token at row dst_row0 + l of its [B, L, C/P]
// buffer.  The destinations are peer GPUs' memory (NVLink P2P stores through symmetric-memory
// mappings), so the exchange costs no pass of its own.  Same arithmetic as rmsnorm_rope_kernel.
struct ScatterDst {
  bf16* ptr[8];
};

__global__ void __launch_bounds__(ROW_THREADS)
rmsnorm_scatter_kernel(const bf16* __restrict__ x, const bf16* __restrict__ weight, ScatterDst dst,
                       int L_local, int C, int group_c, float eps, long long row_stride,
                       long long dst_batch_stride, long long dst_row0) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const int b = static_cast<int>(row / L_local);
  const int l = static_cast<int>(row - static_cast<long long>(b) * L_local);
  const bf16* xr = x + row * row_stride;
  const int nvec = C >> 2;
  float4 v[ROW_MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
    const int vi = threadIdx.x + i * ROW_THREADS;
    if (vi < nvec) {
      v[i] = load4(xr + vi * 4);
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  float rstd = 1.f;
  if (weight != nullptr) rstd = bf16_round(rsqrtf(block_sum(ss, red) / C + eps));
  const long long drow = static_cast<long long>(b) * dst_batch_stride + (dst_row0 + l) * group_c;
#pragma unroll
  for (int i = 0; i < ROW_MAXV; ++i) {
    const int vi = threadIdx.x + i * ROW_THREADS;
    if (vi < nvec) {
      const int c0 = vi * 4;
      float4 y = v[i];
      if (weight != nullptr) {
        const float4 w = load4(weight + c0);
        y.x = bf16_round(bf16_round(y.x * rstd) * w.x);
        y.y = bf16_round(bf16_round(y.y * rstd) * w.y);
        y.z = bf16_round(bf16_round(y.z * rstd) * w.z);
        y.w = bf16_round(bf16_round(y.w * rstd) * w.w);
      }
      const int g = c0 / group_c;                  // group_c % 4 == 0: a vector never straddles groups
      store4(dst.ptr[g] + drow + (c0 - g * group_c), y);
    }
  }
}

}  // namespace m4d

extern "C" int m4d_rmsnorm_scatter(const void* x, long long row_stride, const void* weight, int B,
                                   int L_local, int C, float eps, void* const* dst, int P,
                                   long long dst_batch_stride, long long dst_row0, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && dst && B > 0 && L_local > 0 && C > 0 && P >= 1 && P <= 8, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % (4 * P) == 0 && C <= ROW_THREADS * 4 * ROW_MAXV, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(row_stride >= C && row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0,
              M4D_ERR_ALIGN);
  ScatterDst d;
  for (int g = 0; g < 8; ++g) {
    d.ptr[g] = g < P ? static_cast<bf16*>(dst[g]) : nullptr;
    M4D_REQUIRE(g >= P || (d.ptr[g] != nullptr && (reinterpret_cast<uintptr_t>(d.ptr[g]) & 7) == 0),
                M4D_ERR_ALIGN);
  }
  const long long rows = static_cast<long long>(B) * L_local;
  M4D_REQUIRE(rows < (1ll << 31), M4D_ERR_BAD_SHAPE);
  rmsnorm_scatter_kernel<<<static_cast<unsigned>(rows), ROW_THREADS, 0, stream>>>(
      static_cast<const bf16*>(x), static_cast<const bf16*>(weight), d, L_local, C, C / P, eps, row_stride,
      dst_batch_stride, dst_row0);
  M4D_CHECK_LAUNCH("rmsnorm_scatter_kernel");
  return M4D_OK;
}

extern "C" int m4d_rmsnorm_rope(void* x, long long row_stride, const void* weight,
                                const float* rope_cos, const float* rope_sin, const int* grid_fhw,
                                int B, int L, int heads, int head_dim, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int C = heads * head_dim;
  M4D_REQUIRE(x && B > 0 && L > 0 && heads > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % 4 == 0 && C <= ROW_THREADS * 4 * ROW_MAXV && head_dim % 4 == 0,
              M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(rope_cos == nullptr || grid_fhw != nullptr, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(row_stride >= C && row_stride % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(x) & 7) == 0,
              M4D_ERR_ALIGN);
  const long long rows = static_cast<long long>(B) * L;
  M4D_REQUIRE(rows < (1ll << 31), M4D_ERR_BAD_SHAPE);
  if (head_dim == 128 && C <= 2 * 32 * 8 * RW_MAXV && row_stride % 8 == 0 && aligned16(x) &&
      (weight == nullptr || aligned16(weight))) {
    const unsigned grid = static_cast<unsigned>((rows + RW_WARPS / 2 - 1) / (RW_WARPS / 2));
    auto kern = weight != nullptr ? rmsnorm_rope_warp_kernel<true> : rmsnorm_rope_warp_kernel<false>;
    kern<<<grid, RW_WARPS * 32, 0, stream>>>(
        static_cast<bf16*>(x), static_cast<const bf16*>(weight), rope_cos, rope_sin, grid_fhw, rows, L, C, eps,
        row_stride);
    M4D_CHECK_LAUNCH("rmsnorm_rope_warp_kernel");
    return M4D_OK;
  }
  rmsnorm_rope_kernel<<<static_cast<unsigned>(rows), ROW_THREADS, 0, stream>>>(
      static_cast<bf16*>(x), static_cast<const bf16*>(weight), rope_cos, rope_sin, grid_fhw, L, C,
      head_dim, eps, row_stride);
  M4D_CHECK_LAUNCH("rmsnorm_rope_kernel");
  return M4D_OK;
}

extern "C" int m4d_small_linear_f32(const float* x, const void* w, const void* bias, float* y,
                                    int M, int N, int K, int silu_in, int silu_out,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(M <= 8, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(K % 8 == 0 && aligned16(x) && aligned16(w), M4D_ERR_ALIGN);
  const int grid = (N + 7) / 8;
  if (silu_in)
    small_linear_f32_kernel<1><<<grid, 256, 0, stream>>>(x, static_cast<const bf16*>(w),
                                                         static_cast<const bf16*>(bias), y, M, N, K,
                                                         silu_out);
  else
    small_linear_f32_kernel<0><<<grid, 256, 0, stream>>>(x, static_cast<const bf16*>(w),
                                                         static_cast<const bf16*>(bias), y, M, N, K,
                                                         silu_out);
  M4D_CHECK_LAUNCH("small_linear_f32_kernel");
  return M4D_OK;
}

extern "C" int m4d_timestep_embedding(const float* t, int B, int dim, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(t && out && B > 0 && dim > 0 && dim % 2 == 0, M4D_ERR_BAD_SHAPE);
  const int n = B * (dim / 2);
  timestep_embedding_kernel<<<(n + 127) / 128, 128, 0, stream>>>(t, B, dim, out);
  M4D_CHECK_LAUNCH("timestep_embedding_kernel");
  return M4D_OK;
}

extern "C" int m4d_add_bcast_f32(const void* a_bf16, const float* e, float* out, int B, long long n,
                                 long long m, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(a_bf16 && e && out && B > 0 && n > 0 && m > 0 && n % m == 0, M4D_ERR_BAD_SHAPE);
  const long long total = B * n;
  add_bcast_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      static_cast<const bf16*>(a_bf16), e, out, B, n, m);
  M4D_CHECK_LAUNCH("add_bcast_f32_kernel");
  return M4D_OK;
}

extern "C" int m4d_patchify(const void* x, const void* y, int B, int Cx, int Cy, int T, int H,
                            int W, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && out && B > 0 && Cx > 0 && Cy >= 0 && T > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(H % 2 == 0 && W % 2 == 0 && H > 0 && W > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(Cy == 0 || y != nullptr, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
              M4D_ERR_ALIGN);
  const long long per_b = static_cast<long long>(T) * (H / 2) * (W / 2) * (Cx + Cy);
  dim3 grid(static_cast<unsigned>((per_b + 255) / 256), B);
  patchify_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(x),
                                            static_cast<const bf16*>(y), Cx, Cy, T, H, W,
                                            static_cast<bf16*>(out));
  M4D_CHECK_LAUNCH("patchify_kernel");
  return M4D_OK;
}

extern "C" int m4d_unpatchify(const void* tokens, long long tok_batch_stride, int skip_tokens,
                              int B, int Cout, int T, int H, int W, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(tokens && out && B > 0 && Cout > 0 && T > 0 && H > 0 && W > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(H % 2 == 0 && W % 2 == 0 && skip_tokens >= 0, M4D_ERR_BAD_SHAPE);
  const long long per_b = static_cast<long long>(Cout) * T * H * W;
  dim3 grid(static_cast<unsigned>((per_b + 255) / 256), B);
  unpatchify_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(tokens), tok_batch_stride,
                                              skip_tokens, Cout, T, H, W, static_cast<bf16*>(out));
  M4D_CHECK_LAUNCH("unpatchify_kernel");
  return M4D_OK;
}

extern "C" int m4d_widen_rows(const void* src_bf16, float* dst, int B, int rows, int C,
                              long long dst_batch_stride, int dst_row0, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(src_bf16 && dst && B > 0 && rows > 0 && C > 0 && dst_row0 >= 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % 4 == 0 && aligned16(dst) && dst_batch_stride % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(src_bf16) & 7) == 0,
              M4D_ERR_ALIGN);
  const long long n4 = static_cast<long long>(rows) * (C / 4);
  dim3 grid(static_cast<unsigned>((n4 + 255) / 256), B);
  widen_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(src_bf16), dst, rows, C,
                                              dst_batch_stride, dst_row0);
  M4D_CHECK_LAUNCH("widen_rows_kernel");
  return M4D_OK;
}

extern "C" int m4d_cfg_euler_step(const void* uncond, const void* text, void* latents,
                                  float guidance, float dt, long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(uncond && text && latents && n > 0, M4D_ERR_BAD_SHAPE);
  cfg_euler_step_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<const bf16*>(uncond), static_cast<const bf16*>(text), static_cast<bf16*>(latents),
      guidance, dt, n);
  M4D_CHECK_LAUNCH("cfg_euler_step_kernel");
  return M4D_OK;
}

extern "C" int m4d_silu_bf16(const float* x, void* out, long long n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(x && out && n > 0 && n % 4 == 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(aligned16(x) && (reinterpret_cast<uintptr_t>(out) & 7) == 0, M4D_ERR_ALIGN);
  const long long n4 = n / 4;
  silu_bf16_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(
      x, static_cast<bf16*>(out), n4);
  M4D_CHECK_LAUNCH("silu_bf16_kernel");
  return M4D_OK;
}

#ifdef M4D_DEV
extern "C" void m4d_dev_set_flags(int flags) { g_dev_flags = flags; }
#endif

extern "C" int m4d_version(void) { return 100; }

extern "C" const char* m4d_error_string(int code) {
  switch (code) {
    case M4D_OK: return "ok";
    case M4D_ERR_BAD_SHAPE: return "bad shape or null pointer";
    case M4D_ERR_UNSUPPORTED: return "unsupported configuration (e.g. head_dim != 128)";
    case M4D_ERR_ALIGN: return "misaligned pointer or stride";
    case M4D_ERR_WORKSPACE: return "insufficient workspace";
    case M4D_ERR_CUDA: return "CUDA error (see stderr)";
    case M4D_ERR_NO_DEVICE: return "no CUDA driver / device available";
  }
  return "unknown error";
}

extern "C" int m4d_device_check(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return M4D_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return M4D_ERR_NO_DEVICE;
  if (prop.major != 10) return M4D_ERR_UNSUPPORTED;
  return M4D_OK;
}
