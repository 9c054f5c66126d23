// Motion-Perception-Module front end (wan_transformer4d.py:1127-1156) between the frozen OmniMAE
// trunk (out of scope; its [B, 14*14, 768] patch tokens are the input here) and the per-block
// SpatialGuidanceModule (t4d:739-783, fused into layernorm_modulate):
//
//   feature_adapter = Conv2d(768,768,3,pad 1) -> SiLU -> Conv2d(768,768,3,pad 1)   t4d:888-892,1148
//   F.interpolate(size=(H/2, W/2), mode='bilinear', align_corners=False)            t4d:1149
//   .unsqueeze(2).repeat(1,1,latent_T,1,1).flatten(2).transpose(1,2)                t4d:1150-1151
//
// The trunk's tokens are already channels-last ([B,14,14,768] — the reference permutes them to
// NCHW only for cuDNN), so each 3x3 convolution on the 14x14 map is an im2col gather (this file)
// + the tcgen05 GEMM (gemm.cu) with the tap-major packed weight; 2 x 1.04 GFLOP per sample —
// too small for a dedicated implicit-GEMM kernel.  Everything here is HBM/latency bound.
#include "common.h"
#include "ptx.cuh"

namespace m4d {

__device__ __forceinline__ float mpm_silu(float x) { return x / (1.f + __expf(-x)); }

// rows[(f*H + y)*W + x][tap*C + c] = act(in[f][y+dy-1][x+dx-1][c]) (zero outside), tap = dy*3+dx.
// One thread = 8 channels (16 B) of one (pixel, tap).
__global__ void __launch_bounds__(256)
im2col3x3_cl_kernel(const bf16* __restrict__ in, bf16* __restrict__ rows, int F, int H, int W, int C,
                    int do_silu) {
  const int c8 = C >> 3;
  const long long total = static_cast<long long>(F) * H * W * 9 * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % c8);
    long long r = i / c8;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const int f = static_cast<int>(r / H);
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      v = *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(f) * H + yy) * W + xx) * C + cv * 8);
      if (do_silu) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = mpm_silu(__uint_as_float(w[e] << 16));
          const float b = mpm_silu(__uint_as_float(w[e] & 0xFFFF0000u));
          w[e] = pack_bf16(a, b);
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    *reinterpret_cast<uint4*>(rows + i * 8) = v;
  }
}

// torch's upsample_bilinear2d, align_corners=False, scale derived from the sizes
// (aten area_pixel_compute_source_index): src = max(0, in/out * (dst + 0.5) - 0.5).
__device__ __forceinline__ void bilinear_src(int dst, int in_size, float scale, int& i0, int& i1, float& l1) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

// in [B, h, w, C] -> out [B, T, H, W, C]: bilinear resize (fp32 math, bf16 result) written to all
// T frames; out_silu (optional) receives bf16(SiLU(.)) of the same values — the first op of every
// block's spatial_guide (t4d:746-749), hoisted out of the 2 x num_layers per-block calls.
__global__ void __launch_bounds__(256)
bilinear_repeat_cl_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, bf16* __restrict__ out_silu,
                          int B, int h, int w, int C, int T, int H, int W) {
  const int c8 = C >> 3;
  const long long total = static_cast<long long>(B) * H * W * c8;
  const float sh = static_cast<float>(h) / static_cast<float>(H);
  const float sw = static_cast<float>(w) / static_cast<float>(W);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % c8);
    long long r = i / c8;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(y, h, sh, y0, y1, ly);
    bilinear_src(x, w, sw, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const bf16* base = in + static_cast<long long>(b) * h * w * C + cv * 8;
    const uint4 v00 = *reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * w + x0) * C);
    const uint4 v01 = *reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * w + x1) * C);
    const uint4 v10 = *reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * w + x0) * C);
    const uint4 v11 = *reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * w + x1) * C);
    const uint32_t a[4] = {v00.x, v00.y, v00.z, v00.w}, bq[4] = {v01.x, v01.y, v01.z, v01.w};
    const uint32_t c[4] = {v10.x, v10.y, v10.z, v10.w}, d[4] = {v11.x, v11.y, v11.z, v11.w};
    uint32_t o[4], os[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r2[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float fa = hh ? __uint_as_float(a[e] & 0xFFFF0000u) : __uint_as_float(a[e] << 16);
        const float fb = hh ? __uint_as_float(bq[e] & 0xFFFF0000u) : __uint_as_float(bq[e] << 16);
        const float fc = hh ? __uint_as_float(c[e] & 0xFFFF0000u) : __uint_as_float(c[e] << 16);
        const float fd = hh ? __uint_as_float(d[e] & 0xFFFF0000u) : __uint_as_float(d[e] << 16);
        r2[hh] = hy * (hx * fa + lx * fb) + ly * (hx * fc + lx * fd);
      }
      o[e] = pack_bf16(r2[0], r2[1]);
      os[e] = pack_bf16(mpm_silu(__uint_as_float(o[e] << 16)), mpm_silu(__uint_as_float(o[e] & 0xFFFF0000u)));
    }
    const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]), sv = make_uint4(os[0], os[1], os[2], os[3]);
    for (int t = 0; t < T; ++t) {
      const long long off = (((static_cast<long long>(b) * T + t) * H + y) * W + x) * C + cv * 8;
      if (out != nullptr) *reinterpret_cast<uint4*>(out + off) = ov;
      if (out_silu != nullptr) *reinterpret_cast<uint4*>(out_silu + off) = sv;
    }
  }
}

}  // namespace m4d

using namespace m4d;

static int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

extern "C" int m4d_im2col3x3_cl(const void* in, void* rows, int F, int H, int W, int C, int do_silu,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(in && rows && F > 0 && H > 0 && W > 0 && C > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C % 8 == 0, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(aligned16(in) && aligned16(rows), M4D_ERR_ALIGN);
  const long long total = static_cast<long long>(F) * H * W * 9 * (C / 8);
  im2col3x3_cl_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const bf16*>(in),
                                                           static_cast<bf16*>(rows), F, H, W, C, do_silu);
  M4D_CHECK_LAUNCH("im2col3x3_cl_kernel");
  return M4D_OK;
}

extern "C" int m4d_bilinear_repeat_cl(const void* in, void* out, void* out_silu, int B, int h, int w, int C,
                                      int T, int H, int W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(in && (out || out_silu) && B > 0 && h > 0 && w > 0 && T > 0 && H > 0 && W > 0, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(C > 0 && C % 8 == 0, M4D_ERR_UNSUPPORTED);
  M4D_REQUIRE(aligned16(in) && (out == nullptr || aligned16(out)) && (out_silu == nullptr || aligned16(out_silu)),
              M4D_ERR_ALIGN);
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  bilinear_repeat_cl_kernel<<<grid_for(total), 256, 0, stream>>>(
      static_cast<const bf16*>(in), static_cast<bf16*>(out), static_cast<bf16*>(out_silu), B, h, w, C, T, H, W);
  M4D_CHECK_LAUNCH("bilinear_repeat_cl_kernel");
  return M4D_OK;
}
