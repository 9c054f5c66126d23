// Forward 3D-Gaussian-splatting rasteriser for the stage-1 -> stage-2 hand-off: `render_with_gs`
// (scripts/inference/infer.py:260-273) -> `gs_render` (MoRe4D/utils/gaussian_splatting.py:13-43) ->
// `render_cuda` (:201-281), which calls the third-party `diff_gaussian_rasterization` extension
// (graphdeco-inria @8064f52, README.md:60; NOT in the reference tree) once per frame and camera —
// 11 trajectories x 49 frames of 188 416 gaussians at 368x512 (infer.py:398-444,906-910).
//
// Same published algorithm (Kerbl et al. 2023: per-gaussian projection + EWA covariance, 16x16
// pixel tiles, depth-sorted front-to-back alpha blending with the 1/255 and 1e-4 cut-offs), but laid
// out for this workload instead of being a port:
//   * ALL views of a call (e.g. the 49 frames of one trajectory) go through ONE launch sequence:
//     every kernel's grid spans (view, gaussian) or (view, tile); the reference loops over frames in
//     Python and pays ~10 launches + a device->host sync per frame.
//   * The gaussians are a point cloud with one shared, tiny covariance (scale 1e-4): a splat covers
//     ~1-4 tiles and a tile holds a few hundred splats.  So instead of one global 64-bit radix sort
//     of (tile | depth) keys over all duplicates, splats are BINNED per tile (count -> scan ->
//     scatter) and each tile sorts its own short list by (depth, gaussian index) with a bitonic
//     network in shared memory (global memory for the rare overfull tile).  The order is a total
//     one, so the image is deterministic; ties resolve by gaussian index, which is what the stable
//     radix sort of the original yields.
//   * colours are precomputed (use_sh = False), the background is the wrapper's constant.
// HBM / atomics / shared-memory work; no tensor cores.
#include "common.h"

namespace m4d {

constexpr int GS_TILE = 16;
constexpr int GS_SORT_SMEM = 4096;      // keys sorted in shared memory per tile (32 KB)

struct GsCam {                          // per view, device memory
  float view[12];                       // world -> camera, rows of [R | t]
  float fx, fy, tanx, tany;             // focal lengths in pixels, tan(fov / 2)
  float p00, p11, p22, p23;             // camera -> clip (get_projection_matrix, gaussian_splatting.py:198-226)
};

struct GsGeom {                         // per (view, gaussian)
  float2 xy;
  float depth;
  int radius;                           // 0 = culled
  float4 conic_opacity;
  int4 rect;                            // tile rectangle [x0, y0, x1, y1)
};

__device__ __forceinline__ void gs_rect(float2 p, int r, int gx, int gy, int4& rc) {
  rc.x = min(gx, max(0, static_cast<int>((p.x - r) / GS_TILE)));
  rc.y = min(gy, max(0, static_cast<int>((p.y - r) / GS_TILE)));
  rc.z = min(gx, max(0, static_cast<int>((p.x + r + GS_TILE - 1) / GS_TILE)));
  rc.w = min(gy, max(0, static_cast<int>((p.y + r + GS_TILE - 1) / GS_TILE)));
}

// forward.cu preprocessCUDA + computeCov2D for one (view, gaussian); counts the tiles it touches.
__global__ void __launch_bounds__(256)
gs_preprocess_kernel(const float* __restrict__ means, long long means_vstride, const float* __restrict__ opacity,
                     const float* __restrict__ cov3d, const GsCam* __restrict__ cams, long long N, int V, int H,
                     int W, GsGeom* __restrict__ geom, int* __restrict__ tile_count) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (i >= N) return;
  const GsCam cam = cams[v];
  const float* m = means + v * means_vstride + i * 3;
  const float x = m[0], y = m[1], z = m[2];
  GsGeom g;
  g.radius = 0;
  g.xy = make_float2(0.f, 0.f);
  g.depth = 0.f;
  g.conic_opacity = make_float4(0.f, 0.f, 0.f, 0.f);
  g.rect = make_int4(0, 0, 0, 0);
  const float tx = cam.view[0] * x + cam.view[1] * y + cam.view[2] * z + cam.view[3];
  const float ty = cam.view[4] * x + cam.view[5] * y + cam.view[6] * z + cam.view[7];
  const float tz = cam.view[8] * x + cam.view[9] * y + cam.view[10] * z + cam.view[11];
  GsGeom* out = geom + static_cast<long long>(v) * N + i;
  if (!(tz > 0.2f)) {                                     // in_frustum
    *out = g;
    return;
  }
  // clip = P * p_view (w = z), perspective divide with the extension's + 1e-7
  const float pw = 1.0f / (tz + 0.0000001f);
  const float ndc_x = cam.p00 * tx * pw, ndc_y = cam.p11 * ty * pw;
  // EWA: J from the clamped view-space position, cov2D = (J R) Sigma (J R)^T + 0.3 I
  const float limx = 1.3f * cam.tanx, limy = 1.3f * cam.tany;
  const float cx = fminf(limx, fmaxf(-limx, tx / tz)) * tz;
  const float cy = fminf(limy, fmaxf(-limy, ty / tz)) * tz;
  const float j00 = cam.fx / tz, j02 = -(cam.fx * cx) / (tz * tz);
  const float j11 = cam.fy / tz, j12 = -(cam.fy * cy) / (tz * tz);
  float T0[3], T1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    T0[k] = j00 * cam.view[k] + j02 * cam.view[8 + k];
    T1[k] = j11 * cam.view[4 + k] + j12 * cam.view[8 + k];
  }
  const float sxx = cov3d[0], sxy = cov3d[1], sxz = cov3d[2], syy = cov3d[3], syz = cov3d[4], szz = cov3d[5];
  const float u0 = sxx * T0[0] + sxy * T0[1] + sxz * T0[2];
  const float u1 = sxy * T0[0] + syy * T0[1] + syz * T0[2];
  const float u2 = sxz * T0[0] + syz * T0[1] + szz * T0[2];
  const float w0 = sxx * T1[0] + sxy * T1[1] + sxz * T1[2];
  const float w1 = sxy * T1[0] + syy * T1[1] + syz * T1[2];
  const float w2 = sxz * T1[0] + syz * T1[1] + szz * T1[2];
  const float a = T0[0] * u0 + T0[1] * u1 + T0[2] * u2 + 0.3f;
  const float b = T1[0] * u0 + T1[1] * u1 + T1[2] * u2;
  const float c = T1[0] * w0 + T1[1] * w1 + T1[2] * w2 + 0.3f;
  const float det = a * c - b * b;
  if (det == 0.0f) {
    *out = g;
    return;
  }
  const float inv = 1.0f / det;
  const float mid = 0.5f * (a + c);
  const float root = sqrtf(fmaxf(0.1f, mid * mid - det));
  const int radius = static_cast<int>(ceilf(3.0f * sqrtf(fmaxf(mid + root, mid - root))));
  const float2 p = make_float2(((ndc_x + 1.0f) * W - 1.0f) * 0.5f, ((ndc_y + 1.0f) * H - 1.0f) * 0.5f);
  const int gx = (W + GS_TILE - 1) / GS_TILE, gy = (H + GS_TILE - 1) / GS_TILE;
  int4 rc;
  gs_rect(p, radius, gx, gy, rc);
  if ((rc.z - rc.x) * (rc.w - rc.y) <= 0 || radius <= 0) {
    *out = g;
    return;
  }
  g.radius = radius;
  g.xy = p;
  g.depth = tz;
  g.conic_opacity = make_float4(c * inv, -b * inv, a * inv, opacity[i]);
  g.rect = rc;
  *out = g;
  int* tc = tile_count + static_cast<long long>(v) * gx * gy;
  for (int yy = rc.y; yy < rc.w; ++yy)
    for (int xx = rc.x; xx < rc.z; ++xx) atomicAdd(tc + yy * gx + xx, 1);
}

// exclusive scan of `n` counts by ONE block (n = views x tiles, a few 1e5 at most): offsets[0..n], and
// a copy into `cursor` for the scatter pass
__global__ void __launch_bounds__(1024)
gs_scan_kernel(const int* __restrict__ count, long long n, long long* __restrict__ offsets,
               long long* __restrict__ cursor) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const long long per = (n + 1023) / 1024;
  const long long lo = min(n, t * per), hi = min(n, lo + per);
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += count[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const long long add = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += add;
    __syncthreads();
  }
  long long run = part[t] - s;
  for (long long i = lo; i < hi; ++i) {
    offsets[i] = run;
    cursor[i] = run;
    run += count[i];
  }
  if (t == 1023) offsets[n] = part[1023];
}

// one (depth | gaussian) key per touched tile, appended to the tile's list in arrival order
__global__ void __launch_bounds__(256)
gs_scatter_kernel(const GsGeom* __restrict__ geom, long long N, int gx, int gy, long long* __restrict__ cursor,
                  unsigned long long* __restrict__ keys, long long capacity) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (i >= N) return;
  const GsGeom g = geom[static_cast<long long>(v) * N + i];
  if (g.radius <= 0) return;
  const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(g.depth)) << 32) |
                                 static_cast<unsigned long long>(static_cast<unsigned int>(i));
  long long* cur = cursor + static_cast<long long>(v) * gx * gy;
  for (int yy = g.rect.y; yy < g.rect.w; ++yy)
    for (int xx = g.rect.x; xx < g.rect.z; ++xx) {
      const long long pos = atomicAdd(reinterpret_cast<unsigned long long*>(cur + yy * gx + xx), 1ull);
      if (pos < capacity) keys[pos] = key;
    }
}

// per tile: ascending bitonic sort of the (depth | gaussian index) keys
__global__ void __launch_bounds__(256)
gs_sort_tiles_kernel(const long long* __restrict__ offsets, unsigned long long* __restrict__ keys) {
  __shared__ unsigned long long sk[GS_SORT_SMEM];
  const long long tile = blockIdx.x;
  const long long lo = offsets[tile];
  const int n = static_cast<int>(offsets[tile + 1] - lo);
  if (n <= 1) return;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  unsigned long long* k = keys + lo;
  const bool in_smem = np2 <= GS_SORT_SMEM;
  if (in_smem) {
    for (int i = threadIdx.x; i < np2; i += blockDim.x) sk[i] = i < n ? k[i] : ~0ull;
    __syncthreads();
  }
  // ascending-only formulation (first step of every merge compares mirrored positions, the rest are
  // half-cleaners): with the +inf padding behind the real keys no comparison ever has to move a
  // padding element forward, so on the global-memory path positions >= n are never touched
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const bool flip = stride == (size >> 1);
      for (int i = threadIdx.x; i < (np2 >> 1); i += blockDim.x) {
        int a, b;
        if (flip) {
          const int blk = i / stride, j = i - blk * stride;
          a = blk * size + j;
          b = blk * size + size - 1 - j;
        } else {
          a = 2 * i - (i & (stride - 1));
          b = a + stride;
        }
        if (in_smem) {
          const unsigned long long ka = sk[a], kb = sk[b];
          if (ka > kb) {
            sk[a] = kb;
            sk[b] = ka;
          }
        } else if (b < n) {                                 // overfull tile: same network on global memory
          const unsigned long long ka = k[a], kb = k[b];
          if (ka > kb) {
            k[a] = kb;
            k[b] = ka;
          }
        }
      }
      __syncthreads();
    }
  }
  if (in_smem)
    for (int i = threadIdx.x; i < n; i += blockDim.x) k[i] = sk[i];
}

// forward.cu renderCUDA: one 16x16 block per (view, tile), front-to-back alpha blending
__global__ void __launch_bounds__(GS_TILE* GS_TILE)
gs_render_tiles_kernel(const long long* __restrict__ offsets, const unsigned long long* __restrict__ keys,
                       const GsGeom* __restrict__ geom, const float* __restrict__ colors,
                       long long colors_vstride, long long N, int H, int W, int gx, int gy, float bg0, float bg1,
                       float bg2, float* __restrict__ image) {
  __shared__ float2 s_xy[256];
  __shared__ float4 s_co[256];
  __shared__ float s_rgb[256][3];
  const int tile = blockIdx.x, v = blockIdx.y;
  const int tx = tile % gx, ty = tile / gx;
  const int px = tx * GS_TILE + (threadIdx.x & 15), py = ty * GS_TILE + (threadIdx.x >> 4);
  const bool inside = px < W && py < H;
  const long long lo = offsets[static_cast<long long>(v) * gx * gy + tile];
  const long long hi = offsets[static_cast<long long>(v) * gx * gy + tile + 1];
  const GsGeom* gv = geom + static_cast<long long>(v) * N;
  const float* cv = colors + v * colors_vstride;
  const float pxf = static_cast<float>(px), pyf = static_cast<float>(py);
  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  bool done = !inside;
  for (long long base = lo; base < hi; base += 256) {
    if (__syncthreads_and(done)) break;
    const long long idx = base + threadIdx.x;
    if (idx < hi) {
      const unsigned int gi = static_cast<unsigned int>(keys[idx] & 0xFFFFFFFFull);
      const GsGeom g = gv[gi];
      s_xy[threadIdx.x] = g.xy;
      s_co[threadIdx.x] = g.conic_opacity;
      s_rgb[threadIdx.x][0] = cv[static_cast<long long>(gi) * 3 + 0];
      s_rgb[threadIdx.x][1] = cv[static_cast<long long>(gi) * 3 + 1];
      s_rgb[threadIdx.x][2] = cv[static_cast<long long>(gi) * 3 + 2];
    }
    __syncthreads();
    const int cnt = static_cast<int>(min(256ll, hi - base));
    for (int j = 0; !done && j < cnt; ++j) {
      const float2 xy = s_xy[j];
      const float4 co = s_co[j];
      const float dx = xy.x - pxf, dy = xy.y - pyf;
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.0f) continue;
      const float alpha = fminf(0.99f, co.w * expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = T * (1.0f - alpha);
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const float w = alpha * T;
      C0 += s_rgb[j][0] * w;
      C1 += s_rgb[j][1] * w;
      C2 += s_rgb[j][2] * w;
      T = test_T;
    }
  }
  if (inside) {
    const long long plane = static_cast<long long>(H) * W;
    float* o = image + static_cast<long long>(v) * 3 * plane + static_cast<long long>(py) * W + px;
    o[0] = C0 + T * bg0;
    o[plane] = C1 + T * bg1;
    o[2 * plane] = C2 + T * bg2;
  }
}

// float image [3, H, W] in 0..1 -> uint8 [H, W, 3] with the reference's `* 255` + truncation
// (scripts/inference/infer.py:272-273)
__global__ void gs_to_uint8_kernel(const float* __restrict__ img, unsigned char* __restrict__ out, long long HW,
                                   long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long v = i / (HW * 3), r = i - v * HW * 3;
  const long long p = r / 3;
  const int c = static_cast<int>(r - p * 3);
  const float f = img[(v * 3 + c) * HW + p] * 255.0f;
  out[i] = static_cast<unsigned char>(static_cast<int>(f));
}

static inline long long al256(long long b) { return (b + 255) / 256 * 256; }

}  // namespace m4d

using namespace m4d;

extern "C" long long m4d_gs_render_workspace(long long N, int V, int H, int W, long long dup_capacity) {
  if (N <= 0 || V <= 0 || H <= 0 || W <= 0 || dup_capacity < 0) return -1;
  const long long tiles = static_cast<long long>((W + GS_TILE - 1) / GS_TILE) * ((H + GS_TILE - 1) / GS_TILE) * V;
  return al256(static_cast<long long>(sizeof(GsGeom)) * N * V) + al256(4 * tiles) + 2 * al256(8 * (tiles + 1)) +
         al256(8 * dup_capacity);
}

extern "C" int m4d_gs_render(const float* means, long long means_view_stride, const float* colors,
                             long long colors_view_stride, const float* opacity, const float* cov3d,
                             const float* cams, long long N, int V, int H, int W, float bg0, float bg1, float bg2,
                             float* image, unsigned char* image_u8, void* workspace, long long workspace_bytes,
                             long long dup_capacity, long long* dup_needed, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  M4D_REQUIRE(means && colors && opacity && cov3d && cams && image && workspace, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(N > 0 && N < (1ll << 32) && V > 0 && V <= 65535 && H > 0 && W > 0 && dup_capacity > 0,
              M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(aligned16(workspace) && aligned16(cams), M4D_ERR_ALIGN);
  M4D_REQUIRE(workspace_bytes >= m4d_gs_render_workspace(N, V, H, W, dup_capacity), M4D_ERR_WORKSPACE);
  const int gx = (W + GS_TILE - 1) / GS_TILE, gy = (H + GS_TILE - 1) / GS_TILE;
  const long long tiles = static_cast<long long>(gx) * gy * V;
  M4D_REQUIRE(tiles < (1ll << 31), M4D_ERR_BAD_SHAPE);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  GsGeom* geom = reinterpret_cast<GsGeom*>(ws);
  ws += al256(static_cast<long long>(sizeof(GsGeom)) * N * V);
  int* count = reinterpret_cast<int*>(ws);
  ws += al256(4 * tiles);
  long long* offsets = reinterpret_cast<long long*>(ws);
  ws += al256(8 * (tiles + 1));
  long long* cursor = reinterpret_cast<long long*>(ws);
  ws += al256(8 * (tiles + 1));
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);

  int rc = cuda_ok(cudaMemsetAsync(count, 0, 4 * tiles, stream), "cudaMemsetAsync(gs counts)");
  if (rc != M4D_OK) return rc;
  const dim3 ggrid(static_cast<unsigned>((N + 255) / 256), V);
  gs_preprocess_kernel<<<ggrid, 256, 0, stream>>>(means, means_view_stride, opacity, cov3d,
                                                  reinterpret_cast<const GsCam*>(cams), N, V, H, W, geom, count);
  M4D_CHECK_LAUNCH("gs_preprocess_kernel");
  gs_scan_kernel<<<1, 1024, 0, stream>>>(count, tiles, offsets, cursor);
  M4D_CHECK_LAUNCH("gs_scan_kernel");
  // the number of (tile, gaussian) pairs decides whether the caller's key buffer is large enough —
  // the one device->host read of the call (the original extension does the same to size its buffers)
  long long total = 0;
  rc = cuda_ok(cudaMemcpyAsync(&total, offsets + tiles, 8, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync(gs total)");
  if (rc != M4D_OK) return rc;
  rc = cuda_ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize(gs)");
  if (rc != M4D_OK) return rc;
  if (dup_needed) *dup_needed = total;
  if (total > dup_capacity) return M4D_ERR_WORKSPACE;
  gs_scatter_kernel<<<ggrid, 256, 0, stream>>>(geom, N, gx, gy, cursor, keys, dup_capacity);
  M4D_CHECK_LAUNCH("gs_scatter_kernel");
  gs_sort_tiles_kernel<<<static_cast<unsigned>(tiles), 256, 0, stream>>>(offsets, keys);
  M4D_CHECK_LAUNCH("gs_sort_tiles_kernel");
  gs_render_tiles_kernel<<<dim3(gx * gy, V), GS_TILE * GS_TILE, 0, stream>>>(
      offsets, keys, geom, colors, colors_view_stride, N, H, W, gx, gy, bg0, bg1, bg2, image);
  M4D_CHECK_LAUNCH("gs_render_tiles_kernel");
  if (image_u8 != nullptr) {
    const long long total_px = static_cast<long long>(V) * H * W * 3;
    gs_to_uint8_kernel<<<static_cast<unsigned>((total_px + 255) / 256), 256, 0, stream>>>(
        image, image_u8, static_cast<long long>(H) * W, total_px);
    M4D_CHECK_LAUNCH("gs_to_uint8_kernel");
  }
  return M4D_OK;
}
