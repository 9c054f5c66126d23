// Point-cloud projection with a z-buffer — the hand-off from 4D-STraG (decoded trajectories) to
// 4D-ViSM (rendered novel views): `render_with_project`, scripts/inference/infer.py:222-258, which
// the reference runs as ~25 eager torch ops per frame (project, boolean-mask gathers,
// torch.unique + index_reduce_('amin') for the z-buffer, torch_scatter mean, pad, transpose,
// .cpu().numpy().astype(uint8)).  SURVEY.md §8(f) rank 2.
//
// Three HBM/atomic-bound passes over N = H*W points (one thread per point / pixel, coalesced):
//   1. project: cam = E_inv . [p, 1]; uv = K . cam / (cam.z + eps)  (MoRe4D/utils/project_utils.py:
//      47-71, float32, products and sums in the order written there, no FMA contraction);
//      in-frustum test and the COLUMN-major pixel index floor(u W) * H + floor(v H) exactly as
//      infer.py:228-236; atomicMin of the depth's bit pattern (depth >= 0, so uint order = float
//      order) into the z-buffer; index and depth bits are kept for pass 2.
//   2. accumulate: points whose depth EQUALS the pixel's minimum (ties included, infer.py:241)
//      add their colour and a count with atomicAdd — colours are integer-valued (uint8 images),
//      so the float sums are exact and order-independent.
//   3. finalise: mean colour, [W, H, 3] -> [H, W, 3] transpose, truncation to uint8, and the
//      hole mask (all three channels zero), infer.py:246-256.
#include <vector>

#include "common.h"

namespace m4d {

struct ProjParams {
  float e[12];      // first three rows of the world->camera matrix
  float k[6];       // first two rows of the intrinsic matrix
};

// blockIdx.y = view.  `cams` (device, [V][18] = ProjParams per view) replaces the by-value camera
// when several views are rendered in one launch sequence; per-view strides of 0 share the points.
__global__ void __launch_bounds__(256)
project_zmin_kernel(const float* __restrict__ pts, long long pts_view_stride, ProjParams pr,
                    const ProjParams* __restrict__ cams, long long N, int H, int W,
                    int* __restrict__ idx_out, unsigned* __restrict__ dbits_out, unsigned* __restrict__ zbuf) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int view = blockIdx.y;
  if (cams != nullptr) pr = cams[view];
  pts += view * pts_view_stride;
  idx_out += view * N;
  dbits_out += view * N;
  zbuf += static_cast<long long>(view) * H * W;
  const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  float cam[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    cam[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pr.e[4 * r], x), __fmul_rn(pr.e[4 * r + 1], y)),
                                 __fmul_rn(pr.e[4 * r + 2], z)),
                       pr.e[4 * r + 3]);
  const float depth = cam[2];
  const float den = __fadd_rn(depth, 1.1920928955078125e-07f);
  float p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float q = __fdiv_rn(cam[r], den);
    if (isnan(q)) q = 0.f;                              // nan_to_num (project_utils.py:54)
    else if (isinf(q)) q = q > 0.f ? 1e8f : -1e8f;
    p[r] = q;
  }
  const float u = __fadd_rn(__fadd_rn(__fmul_rn(pr.k[0], p[0]), __fmul_rn(pr.k[1], p[1])), __fmul_rn(pr.k[2], p[2]));
  const float v = __fadd_rn(__fadd_rn(__fmul_rn(pr.k[3], p[0]), __fmul_rn(pr.k[4], p[1])), __fmul_rn(pr.k[5], p[2]));
  int idx = -1;
  unsigned db = 0;
  if (u >= 0.f && u <= 1.f && v >= 0.f && v <= 1.f && depth >= 0.f) {
    const float fx = fminf(fmaxf(floorf(__fmul_rn(u, static_cast<float>(W))), 0.f), static_cast<float>(W - 1));
    const float fy = fminf(fmaxf(floorf(__fmul_rn(v, static_cast<float>(H))), 0.f), static_cast<float>(H - 1));
    idx = static_cast<int>(__fadd_rn(__fmul_rn(fx, static_cast<float>(H)), fy));
    db = __float_as_uint(__fadd_rn(depth, 0.f));        // -0.0 -> +0.0
    atomicMin(&zbuf[idx], db);
  }
  idx_out[i] = idx;
  dbits_out[i] = db;
}

__global__ void __launch_bounds__(256)
project_accum_kernel(const float* __restrict__ colors, long long colors_view_stride,
                     const int* __restrict__ idx_in, const unsigned* __restrict__ dbits,
                     const unsigned* __restrict__ zbuf, long long N, long long hw, float* __restrict__ acc) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int view = blockIdx.y;
  colors += view * colors_view_stride;
  idx_in += view * N;
  dbits += view * N;
  zbuf += view * hw;
  acc += 4 * view * hw;
  const int idx = idx_in[i];
  if (idx < 0 || dbits[i] != zbuf[idx]) return;
  float* a = acc + 4ll * idx;
  atomicAdd(a + 0, colors[3 * i]);
  atomicAdd(a + 1, colors[3 * i + 1]);
  atomicAdd(a + 2, colors[3 * i + 2]);
  atomicAdd(a + 3, 1.0f);
}

__global__ void __launch_bounds__(256)
project_finalize_kernel(const float* __restrict__ acc, int H, int W, unsigned char* __restrict__ image,
                        unsigned char* __restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // output pixel h * W + w
  if (i >= H * W) return;
  const long long hw = static_cast<long long>(H) * W;
  acc += 4 * blockIdx.y * hw;
  image += 3 * blockIdx.y * hw;
  mask += blockIdx.y * hw;
  const int h = i / W, w = i - h * W;
  const float4 a = reinterpret_cast<const float4*>(acc)[static_cast<long long>(w) * H + h];
  unsigned char c[3] = {0, 0, 0};
  if (a.w > 0.f) {
    c[0] = static_cast<unsigned char>(static_cast<int>(__fdiv_rn(a.x, a.w)));
    c[1] = static_cast<unsigned char>(static_cast<int>(__fdiv_rn(a.y, a.w)));
    c[2] = static_cast<unsigned char>(static_cast<int>(__fdiv_rn(a.z, a.w)));
  }
  image[3ll * i] = c[0];
  image[3ll * i + 1] = c[1];
  image[3ll * i + 2] = c[2];
  mask[i] = (c[0] | c[1] | c[2]) == 0;
}

}  // namespace m4d

using namespace m4d;

static long long project_ws_bytes(long long N, int H, int W, int V) {
  const long long hw = static_cast<long long>(H) * W;
  // float4 accumulators, z-buffer, idx + depth bits (per view), then the V device-side cameras
  return V * (hw * 16 + hw * 4 + N * 8) + (V > 1 ? V * static_cast<long long>(sizeof(ProjParams)) + 16 : 0);
}

static int project_impl(const float* points, long long pts_view_stride, const float* colors,
                        long long colors_view_stride, const float* world2cam, const float* intrinsic,
                        long long N, int V, int H, int W, unsigned char* image, unsigned char* mask,
                        void* workspace, long long workspace_bytes, cudaStream_t stream) {
  M4D_REQUIRE(points && colors && world2cam && intrinsic && image && mask && workspace, M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(N > 0 && H > 0 && W > 0 && V > 0 && V <= 65535 && static_cast<long long>(H) * W < (1ll << 24),
              M4D_ERR_BAD_SHAPE);
  M4D_REQUIRE(workspace_bytes >= project_ws_bytes(N, H, W, V), M4D_ERR_WORKSPACE);
  M4D_REQUIRE(aligned16(workspace), M4D_ERR_ALIGN);
  const long long hw = static_cast<long long>(H) * W;
  float* acc = static_cast<float*>(workspace);                                  // [V][hw][4], 16-byte aligned
  unsigned* zbuf = reinterpret_cast<unsigned*>(acc + 4 * hw * V);
  int* idx = reinterpret_cast<int*>(zbuf + hw * V);
  unsigned* dbits = reinterpret_cast<unsigned*>(idx + N * V);
  // the 4x4 / 3x3 matrices are HOST pointers (16 + 9 floats per view): one view travels as a kernel
  // argument, several are staged into the workspace with one async copy
  ProjParams pr = {};
  const ProjParams* cams = nullptr;
  if (V == 1) {
    for (int i = 0; i < 12; ++i) pr.e[i] = world2cam[i];
    for (int i = 0; i < 6; ++i) pr.k[i] = intrinsic[i];
  } else {
    static_assert(sizeof(ProjParams) == 72, "ProjParams is 18 floats");
    std::vector<ProjParams> host(V);
    for (int v = 0; v < V; ++v) {
      for (int i = 0; i < 12; ++i) host[v].e[i] = world2cam[16 * v + i];
      for (int i = 0; i < 6; ++i) host[v].k[i] = intrinsic[i];
    }
    uintptr_t cp = (reinterpret_cast<uintptr_t>(dbits + N * V) + 15) & ~static_cast<uintptr_t>(15);
    cams = reinterpret_cast<const ProjParams*>(cp);
    // pageable source: the copy is staged by the runtime before the call returns
    int rc = cuda_ok(cudaMemcpyAsync(reinterpret_cast<void*>(cp), host.data(), V * sizeof(ProjParams),
                                     cudaMemcpyHostToDevice, stream), "memcpy(project cameras)");
    if (rc != M4D_OK) return rc;
  }
  int rc = cuda_ok(cudaMemsetAsync(acc, 0, hw * 16 * V, stream), "memset(project acc)");
  if (rc != M4D_OK) return rc;
  rc = cuda_ok(cudaMemsetAsync(zbuf, 0xFF, hw * 4 * V, stream), "memset(project zbuf)");
  if (rc != M4D_OK) return rc;
  const dim3 nb(static_cast<unsigned>((N + 255) / 256), V);
  project_zmin_kernel<<<nb, 256, 0, stream>>>(points, pts_view_stride, pr, cams, N, H, W, idx, dbits, zbuf);
  M4D_CHECK_LAUNCH("project_zmin_kernel");
  project_accum_kernel<<<nb, 256, 0, stream>>>(colors, colors_view_stride, idx, dbits, zbuf, N, hw, acc);
  M4D_CHECK_LAUNCH("project_accum_kernel");
  project_finalize_kernel<<<dim3(static_cast<unsigned>((hw + 255) / 256), V), 256, 0, stream>>>(acc, H, W, image, mask);
  M4D_CHECK_LAUNCH("project_finalize_kernel");
  return M4D_OK;
}

extern "C" long long m4d_project_points_workspace(long long N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return project_ws_bytes(N, H, W, 1);
}

extern "C" int m4d_project_points(const float* points, const float* colors, const float* world2cam,
                                  const float* intrinsic, long long N, int H, int W, unsigned char* image,
                                  unsigned char* mask, void* workspace, long long workspace_bytes,
                                  void* stream_) {
  return project_impl(points, 0, colors, 0, world2cam, intrinsic, N, 1, H, W, image, mask, workspace,
                      workspace_bytes, static_cast<cudaStream_t>(stream_));
}

extern "C" long long m4d_project_views_workspace(long long N, int V, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0 || V <= 0) return 0;
  return project_ws_bytes(N, H, W, V);
}

extern "C" int m4d_project_views(const float* points, long long points_view_stride, const float* colors,
                                 long long colors_view_stride, const float* world2cam, const float* intrinsic,
                                 long long N, int V, int H, int W, unsigned char* image, unsigned char* mask,
                                 void* workspace, long long workspace_bytes, void* stream_) {
  return project_impl(points, points_view_stride, colors, colors_view_stride, world2cam, intrinsic, N, V, H, W,
                      image, mask, workspace, workspace_bytes, static_cast<cudaStream_t>(stream_));
}
