// Shared by the two implicit-GEMM convolution kernels (conv.cu: one TMA box per tap;
// conv_halo.cu: halo tile staged once per time tap): launch parameters and the epilogue of one
// 32-channel chunk of one output pixel.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace m4d {

struct ConvParams {
  int T_out, H_out, W_out;
  int Cin, Cout;            // Cin multiple of 32 (as stored), Cout = real output channels
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int ksub;                 // conv.cu: 32-channel boxes per pipeline stage
  int NT;                   // output-channel tile (multiple of 16, <= 256)
  int n_tiles;              // ceil(Cout / NT)
  int stages;               // conv.cu: pipeline stages; conv_halo.cu: weight-ring stages
  // output addressing: element (t, h, w, n) goes to frame t*t_mul + t_off + n / n_split,
  // channel n % n_split of a channels-last tensor with out_C channels per pixel
  void* out;
  const bf16* residual;     // same addressing as out (NHWC mode only), or null
  const bf16* bias;         // [Cout] or null
  int out_C, t_mul, t_off, n_split;
  int out_mode;             // 0 NHWC bf16; 1 planar NCTHW bf16 (n < Cout)
  int act;                  // 0 none, 1 clamp(-1,1), 2 sigmoid(y + skip)
  const bf16* skip;         // planar NCTHW tensor added before the sigmoid (act == 2)
  long long planar_cstride; // T*H*W of the planar tensors
  // conv_halo.cu only
  int a_stages;             // halo-tile ring depth
  int acc_bufs;             // 1 or 2 accumulator sets in TMEM
  int acc_stride;           // TMEM columns per sub-tile accumulator (NT rounded up to 32)
  int desc_mode;            // unused (kept 0): see the base-offset note in conv_halo.cu
  // fused RMS_norm (+ SiLU) of the conv output (conv_halo.cu, NT == Cout in {96, 192}):
  // norm_out = [silu](rms_norm(y) * gamma), same NHWC addressing as `out`; `out` may be null
  // when only the normalised tensor is consumed (ResidualBlock conv1 -> norm2, vae:198-202)
  const bf16* norm_gamma;
  bf16* norm_out;
  int norm_silu;
  int vec_ok;               // NHWC rows, bias and residual allow 16-byte vector access per 32 channels;
                            // 2: rows (out, residual, norm_out) are also 32-byte aligned -> 256-bit accesses
  // fused GroupNorm statistics (conv_halo.cu, NT == Cout == 128, 32 groups of 4 channels): per
  // spatial tile the sums and sums of squares of the bf16 outputs, [tile][group][2] floats
  float* gn_partials;
};


__device__ __forceinline__ void unpack8(const uint4 u, float* f) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xFFFF0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xFFFF0000u);
}

// Final bf16 values (as floats) of 32 consecutive, fully valid output channels of one pixel:
// bias (16-byte aligned vector loads), bf16 rounding (the reference's bf16 conv output), then
// the residual add (ResidualBlock, vae:224) and its rounding.  `bias`, `res` point at the
// chunk's first channel or are null.
__device__ __forceinline__ void load_res_chunk(const bf16* res, uint4* r4) {
#pragma unroll
  for (int q = 0; q < 4; ++q) r4[q] = reinterpret_cast<const uint4*>(res)[q];
}
// the same 64 bytes as two 256-bit loads (32-byte aligned rows: ConvParams::vec_ok == 2)
__device__ __forceinline__ void load_res_chunk32(const bf16* res, uint4* r4) {
  ldg256(res, reinterpret_cast<uint32_t*>(r4));
  ldg256(res + 16, reinterpret_cast<uint32_t*>(r4 + 2));
}

// `res4`: the chunk's 32 residual values, already in registers (loaded one chunk ahead so their
// latency overlaps the previous chunk / the wait for the accumulators), or null.
__device__ __forceinline__ void conv_chunk_values(const uint32_t* rr, const bf16* bias, const uint4* res4,
                                                  float* v) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
  if (bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float b8[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + q), b8);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[q * 8 + e] += b8[e];
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
  if (res4 != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float r8[8];
      unpack8(res4[q], r8);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[q * 8 + e] = bf16_round(v[q * 8 + e] + r8[e]);
    }
  }
}

// Packed form of conv_chunk_values: 16 words of two bf16 channels each.  bias add in fp32 and one
// rounding (the bf16 conv output), residual add as add.rn.bf16x2 (== fp32 add + rounding).
template <int Q>                                         // Q 16-byte groups = 8 * Q channels
__device__ __forceinline__ void conv_packed(const uint32_t* rr, const bf16* bias, const uint4* res4,
                                            uint32_t* y) {
  if (bias != nullptr) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(bias) + q);
      const uint32_t bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        y[q * 4 + e] = pack_bf16(__uint_as_float(rr[q * 8 + 2 * e]) + bf16_lo(bw[e]),
                                 __uint_as_float(rr[q * 8 + 2 * e + 1]) + bf16_hi(bw[e]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4 * Q; ++i) y[i] = pack_bf16(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
  }
  if (res4 != nullptr) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      y[q * 4 + 0] = add_bf16x2(y[q * 4 + 0], res4[q].x);
      y[q * 4 + 1] = add_bf16x2(y[q * 4 + 1], res4[q].y);
      y[q * 4 + 2] = add_bf16x2(y[q * 4 + 2], res4[q].z);
      y[q * 4 + 3] = add_bf16x2(y[q * 4 + 3], res4[q].w);
    }
  }
}

__device__ __forceinline__ void conv_chunk_packed(const uint32_t* rr, const bf16* bias, const uint4* res4,
                                                  uint32_t* y) {
  conv_packed<4>(rr, bias, res4, y);
}

__device__ __forceinline__ void store_chunk_packed32(bf16* o, const uint32_t* y) {
  stg256(o, y);
  stg256(o + 16, y + 8);
}

__device__ __forceinline__ void store_chunk_packed(bf16* o, const uint32_t* y) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    reinterpret_cast<uint4*>(o)[q] = make_uint4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
}

__device__ __forceinline__ void store_chunk_bf16(bf16* o, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
    reinterpret_cast<uint4*>(o)[q] = u;
  }
}

// Ragged / planar tail of the epilogue (channel counts that are not multiples of 32, unaligned
// rows, the 3-channel planar outputs with clamp / sigmoid): element-wise, kept out of line so
// its dynamically indexed accumulators do not drag the common path into local memory.
static __device__ __noinline__ void conv_store_chunk_slow(const ConvParams& p, const uint32_t* rr, int t, int h,
                                                   int w, int n0) {
  const int nvalid = min(32, p.Cout - n0);
  if (p.out_mode == 0) {
    const int fo = t * p.t_mul + p.t_off + n0 / p.n_split;
    const int ch = n0 % p.n_split;
    const long long off = ((static_cast<long long>(fo) * p.H_out + h) * p.W_out + w) * p.out_C + ch;
    bf16* o = reinterpret_cast<bf16*>(p.out) + off;
    for (int i = 0; i < nvalid; ++i) {
      float y = __uint_as_float(rr[i]);
      if (p.bias != nullptr) y += __bfloat162float(p.bias[n0 + i]);
      y = bf16_round(y);
      if (p.residual != nullptr) y += __bfloat162float(p.residual[off + i]);
      o[i] = __float2bfloat16_rn(y);
    }
  } else {
    // planar NCTHW output of a few channels (decoder head / adaptor conv_out)
    const long long pix = (static_cast<long long>(t) * p.H_out + h) * p.W_out + w;
    for (int i = 0; i < nvalid; ++i) {
      float y = __uint_as_float(rr[i]);
      if (p.bias != nullptr) y += __bfloat162float(p.bias[n0 + i]);
      y = bf16_round(y);
      const long long off = (n0 + i) * p.planar_cstride + pix;
      if (p.act == 1) y = fminf(1.f, fmaxf(-1.f, y));
      if (p.act == 2) {
        y = bf16_round(y + __bfloat162float(p.skip[off]));
        y = 1.f / (1.f + __expf(-y));
      }
      reinterpret_cast<bf16*>(p.out)[off] = __float2bfloat16_rn(y);
    }
  }
}

// Epilogue of output pixel (t, h, w), channels [n0, n0 + 32): rr = fp32 accumulators.
// bias, bf16 rounding (the reference's bf16 conv output), residual add, store.
__device__ __forceinline__ void conv_store_chunk(const ConvParams& p, const uint32_t* rr, int t, int h,
                                                 int w, int n0) {
  if (p.out_mode == 0 && n0 + 32 <= p.Cout && p.vec_ok) {
    const int fo = t * p.t_mul + p.t_off + n0 / p.n_split;
    const int ch = n0 % p.n_split;
    const long long off = ((static_cast<long long>(fo) * p.H_out + h) * p.W_out + w) * p.out_C + ch;
    uint4 r4[4];
    uint32_t y[16];
    const bool a32 = ((reinterpret_cast<uintptr_t>(p.out) + 2 * off) & 31) == 0;    // whole 32-byte sectors
    if (p.residual != nullptr) load_res_chunk(p.residual + off, r4);
    conv_chunk_packed(rr, p.bias ? p.bias + n0 : nullptr, p.residual ? r4 : nullptr, y);
    if (a32) store_chunk_packed32(reinterpret_cast<bf16*>(p.out) + off, y);
    else store_chunk_packed(reinterpret_cast<bf16*>(p.out) + off, y);
  } else {
    conv_store_chunk_slow(p, rr, t, h, w, n0);
  }
}

// conv_halo.cu
bool conv_halo_eligible(int Cin, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                        int T_in, int H_in, int W_in, int T_out, int H_out, int W_out);
int conv_halo_launch(const void* x, int T_in, int H_in, int W_in, const void* w_packed, int Cout_pad,
                     ConvParams p, cudaStream_t stream);

}  // namespace m4d
