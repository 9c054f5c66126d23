/* more4d_b200 — C ABI of libmore4d_sm100.so
 *
 * Drop-in boundary for the 4D-STraG denoising hot path of Zhangyr2022/MoRe4D on NVIDIA B200
 * (sm_100a).  The reference has no FFI of its own: its hot path is Python calling torch /
 * flash-attn library kernels.  Each entry point below replaces the library
 * call(s) cited next to it; the Python shim in more4d_b200/ binds them with ctypes and
 * mirrors the reference's call surface.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all tensor pointers are DEVICE pointers
 *   - "bf16" = __nv_bfloat16 storage; strides are in ELEMENTS unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream)
 *   - the caller owns every buffer; the library never allocates or keeps device pointers
 *   - return 0 on success, a negative M4D_ERR_* otherwise (m4d_error_string decodes it);
 *     launches are asynchronous, so device-side faults surface at the caller's next sync
 *   - no CPU fallback: without an sm_100 device every compute entry point fails
 */
#ifndef MORE4D_B200_H
#define MORE4D_B200_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
  M4D_OK = 0,
  M4D_ERR_BAD_SHAPE = -1,
  M4D_ERR_UNSUPPORTED = -2,
  M4D_ERR_ALIGN = -3,
  M4D_ERR_WORKSPACE = -4,
  M4D_ERR_CUDA = -5,
  M4D_ERR_NO_DEVICE = -6
};

/* GEMM epilogues.  Every one first rounds (acc + bias) to bf16, which is what the reference's
 * autocast nn.Linear hands to the next op. */
enum {
  M4D_EPI_BF16 = 0,              /* out bf16 = acc + bias                                   */
  M4D_EPI_GELU_TANH = 1,         /* out bf16 = GELU_tanh(.)   ffn.1, text_embedding.1       */
  M4D_EPI_GELU_ERF = 2,          /* out bf16 = GELU_erf(.)    img_emb.proj.2                */
  M4D_EPI_F32 = 3,               /* out fp32 = (.)            patch/ref embedding -> x      */
  M4D_EPI_GATE_RESIDUAL_F32 = 4, /* out fp32 = residual + (.) * gate[row / rows_per_batch]  */
  M4D_EPI_ADD_BF16 = 5,          /* out bf16 = (.) + residual_bf16   VAE AttentionBlock.proj */
  M4D_EPI_F32_RAW = 6,           /* out fp32 = acc, no bias, no rounding (attention logits)  */
  M4D_EPI_COUNT = 7
};

int m4d_version(void);
const char* m4d_error_string(int code);
/* M4D_OK iff the current device is compute capability 10.x. */
int m4d_device_check(void);

/* out[M,N] = epilogue(a[M,K] . w[N,K]^T + bias[N]) — tcgen05/TMEM GEMM, TMA-fed.
 * Replaces nn.Linear (cuBLAS) at wan_transformer4d.py:446-448,465 (self-attn q/k/v/o),
 * :481-483,527-531,553 (cross-attn), :620-622 (FFN), :720 (head), :900-902 (text embedding),
 * :729-733 (MLPProj) and, fed by m4d_patchify, the patch/ref convolutions :1073,1087.
 * a, w: bf16 row-major with row strides lda, ldw (multiples of 8); bias: bf16 or NULL.
 * M4D_EPI_GATE_RESIDUAL_F32 fuses `x = x + y * e[k]` (:669,:684) / `x = x + y` (:674):
 * residual fp32 [M, ldr] (may alias out), gate fp32 [*, N] with batch stride
 * gate_batch_stride, or NULL for gate = 1.  M4D_EPI_ADD_BF16 takes a bf16 residual. */
int m4d_gemm_bf16(const void* a, long long lda, const void* w, long long ldw, const void* bias,
                  void* out, long long ldo, int M, int N, int K, int epilogue,
                  const void* residual, long long ldr, const float* gate,
                  long long gate_batch_stride, int rows_per_batch, void* stream);

/* out[b,l,h,:] = softmax(q k^T * scale) v, non-causal, head_dim 128, bf16 "NHD" layout
 * [B, L, heads, 128] with explicit batch / token strides (so q,k,v may be views).
 * Replaces attention()/flash_attention() wan_transformer4d.py:66-236 (flash_attn_varlen_func
 * :138-169, SDPA :232).  k_lens (device int32 [B] or NULL) trims keys per sample like the
 * varlen path.  softmax_scale <= 0 selects 1/sqrt(128).  accumulate != 0 adds the result into
 * `out` in bf16 (the summed image+text cross-attention, :552). */
int m4d_attention_fwd(const void* q, const void* k, const void* v, void* out, int B, int Lq,
                      int Lk, int heads, int head_dim, long long q_stride_b, long long q_stride_l,
                      long long kv_stride_b, long long kv_stride_l, long long out_stride_b,
                      long long out_stride_l, const int* k_lens, float softmax_scale,
                      int accumulate, void* stream);

/* TWO independently normalised attentions over one query set in ONE launch, summed in bf16:
 *   out = bf16( bf16(softmax(q k[:seg]^T) v[:seg]) + bf16(softmax(q k[seg:]^T) v[seg:]) )
 * — WanI2VCrossAttention, wan_transformer4d.py:533-552 (text context then CLIP-image context;
 * bf16 addition commutes, so the order of the two terms is free).  seg_len must be a positive
 * multiple of 128 and < Lk; bit-identical to m4d_attention_fwd on k[:seg] followed by
 * m4d_attention_fwd(accumulate=1) on k[seg:]. */
int m4d_attention_fwd_seg2(const void* q, const void* k, const void* v, void* out, int B, int Lq,
                           int Lk, int seg_len, int heads, int head_dim, long long q_stride_b,
                           long long q_stride_l, long long kv_stride_b, long long kv_stride_l,
                           long long out_stride_b, long long out_stride_l, float softmax_scale,
                           void* stream);

/* m4d_attention_fwd whose epilogue SCATTERS the query rows over n_out <= 8 output buffers: row l goes
 * to out[l / rows_per_out] (HOST array of DEVICE pointers, possibly peer-GPU memory) at row
 * l % rows_per_out, with the given batch / token strides.  The returning half of the
 * sequence-parallel (Ulysses) exchange — the slot of the reference's usp_attn_forward,
 * wan_transformer4d.py:1038-1044 — fused into the attention epilogue: every rank's token chunk of
 * this rank's heads is stored straight into that rank's buffer over NVLink, in ONE launch. */
int m4d_attention_fwd_scatter(const void* q, const void* k, const void* v, void* const* out, int n_out,
                              int rows_per_out, int B, int Lq, int Lk, int heads, int head_dim,
                              long long q_stride_b, long long q_stride_l, long long kv_stride_b,
                              long long kv_stride_l, long long out_stride_b, long long out_stride_l,
                              const int* k_lens, float softmax_scale, void* stream);

/* out = LayerNorm(x) [* weight + bias] [* (1 + scale[b]) + shift[b]], one pass.
 * Replaces F.layer_norm + modulation at wan_transformer4d.py:662,677 (norm1/norm2 + AdaLN),
 * :674 (norm3, affine), :720 (head), :729,733 (MLPProj).  x fp32 or bf16 [rows, C]; weight,
 * bias bf16 [C] or NULL; shift, scale fp32 [*, C] (batch stride mod_batch_stride) or NULL;
 * out bf16 or fp32.  Optional Motion-Perception-Module injection (:757-783): sg bf16
 * [B, sg_rows, 2C] = Linear(SiLU(features)) as (scale | shift), sg_gate bf16 [C]. */
int m4d_layernorm_modulate(const void* x, int x_is_bf16, const void* weight, const void* bias,
                           const float* shift, const float* scale, long long mod_batch_stride,
                           long long rows, int rows_per_batch, int C, float eps, void* out,
                           int out_is_f32, const void* sg, long long sg_batch_stride, int sg_rows,
                           const void* sg_gate, void* stream);

/* In place on bf16 x [B, L, heads*head_dim] (row stride row_stride): WanRMSNorm over the full
 * channel dim (wan_transformer4d.py:378-394; weight NULL skips it) then 3-axis RoPE on
 * adjacent pairs (:340-375; rope_cos NULL skips it).  rope_cos/rope_sin: fp32 [1024, 64]
 * roundings of the reference's float64 `freqs` table; grid_fhw: device int32 [B, 3]. */
int m4d_rmsnorm_rope(void* x, long long row_stride, const void* weight, const float* rope_cos,
                     const float* rope_sin, const int* grid_fhw, int B, int L, int heads,
                     int head_dim, float eps, void* stream);

/* WanRMSNorm over the full channel dim (wan_transformer4d.py:378-394; weight NULL = plain copy)
 * whose output is scattered by head group: dst[g] (HOST array of P <= 8 DEVICE pointers, possibly
 * peer-GPU memory) receives channels [g*C/P, (g+1)*C/P) of local token (b, l) at element
 * b*dst_batch_stride + (dst_row0 + l)*(C/P).  The sending half of the sequence-parallel
 * (Ulysses) exchange around self-attention — the slot of the reference's usp_attn_forward,
 * wan_transformer4d.py:1038-1044 — fused into the normalisation pass over NVLink P2P stores. */
int m4d_rmsnorm_scatter(const void* x, long long row_stride, const void* weight, int B, int L_local,
                        int C, float eps, void* const* dst, int P, long long dst_batch_stride,
                        long long dst_row0, void* stream);

/* y[M,N] fp32 = [SiLU](x[M,K] fp32) . w[N,K]^T(bf16) + bias, optional SiLU on y; M <= 8.
 * The time-embedding MLPs, which the reference runs in fp32 (wan_transformer4d.py:1160-1171). */
int m4d_small_linear_f32(const float* x, const void* w, const void* bias, float* y, int M, int N,
                         int K, int silu_in, int silu_out, void* stream);

/* sinusoidal_embedding_1d (wan_transformer4d.py:239-249): out fp32 [B, dim] = cos | sin,
 * evaluated in float64. */
int m4d_timestep_embedding(const float* t, int B, int dim, float* out, void* stream);

/* out[b, i] = a[i] + e[b, i % m], i < n, m | n — `modulation + e0` (wan_transformer4d.py:659,
 * m == n) and the head's `modulation + e.unsqueeze(1)` (:717, n == 2m). */
int m4d_add_bcast_f32(const void* a_bf16, const float* e, float* out, int B, long long n,
                      long long m, void* stream);

/* im2col of the (1,2,2) patch convolution over cat(x, y) channels
 * (wan_transformer4d.py:1069-1073): out bf16 [B, T*(H/2)*(W/2), (Cx+Cy)*4]. */
int m4d_patchify(const void* x, const void* y, int B, int Cx, int Cy, int T, int H, int W,
                 void* out, void* stream);

/* unpatchify (wan_transformer4d.py:1343-1366) after dropping `skip_tokens` reference tokens
 * (:1323-1326): tokens bf16 [B, *, 4*Cout] -> out bf16 [B, Cout, T, H, W]. */
int m4d_unpatchify(const void* tokens, long long tok_batch_stride, int skip_tokens, int B,
                   int Cout, int T, int H, int W, void* out, void* stream);

/* dst[b, dst_row0 + r, :] (fp32) = src[b*rows + r, :] (bf16): token rows into the fp32
 * residual stream (wan_transformer4d.py:1099-1106 concat/pad). */
int m4d_widen_rows(const void* src_bf16, float* dst, int B, int rows, int C,
                   long long dst_batch_stride, int dst_row0, void* stream);

/* CFG combine + flow-matching Euler step on bf16 latents
 * (pipeline_wan_fun_control.py:820-825). */
int m4d_cfg_euler_step(const void* uncond, const void* text, void* latents, float guidance,
                       float dt, long long n, void* stream);

/* ---- Motion-Sensitive VAE (wan_vae.py, trajectory_module.py), channels-last activations ---- */

/* Causal 3-D / 2-D convolution as a tcgen05 implicit GEMM over a whole channels-last sequence
 * x bf16 [T_in, H_in, W_in, Cin] (Cin % 32 == 0).  Replaces CausalConv3d.forward
 * (wan_vae.py:32-40: cache cat + F.pad + cuDNN Conv3d), Resample's Conv2d's (:82-100) and the
 * adaptors' Conv2d's (trajectory_module.py:73-87).  Padding (pt frames in FRONT, ph/pw on both
 * spatial sides, right/bottom for strided convs) is TMA out-of-bounds zero fill.
 * w_packed: bf16 [Cout_pad, kt*kh*kw*Cin], tap-major / channel-minor, rows >= Cout zero.
 * Output element (t, h, w, n) goes to frame t*t_mul + t_off + n / n_split, channel n % n_split
 * of a channels-last tensor with out_C channels (upsample3d's channel->frame interleave,
 * wan_vae.py:138-141, is n_split = C); residual: same layout, added after bf16 rounding
 * (ResidualBlock, :224).  out_mode 1 writes planar [Cout, T, H, W] instead, with act 1 =
 * clamp(-1,1) (:827) or act 2 = sigmoid(y + skip) (trajectory_module.py:194). */
int m4d_conv_cl(const void* x, int T_in, int H_in, int W_in, int Cin, const void* w_packed,
                int Cout, int Cout_pad, const void* bias, int kt, int kh, int kw, int st, int sh,
                int sw, int pt, int ph, int pw, int T_out, int H_out, int W_out, void* out,
                int out_C, int t_mul, int t_off, int n_split, const void* residual, int out_mode,
                int act, const void* skip, void* stream);

/* The 3x3 (x kt, causal) stride-1 case of m4d_conv_cl with the CONSUMER's RMS_norm [+ SiLU] fused
 * into the epilogue: ResidualBlock = [RMS_norm, SiLU, conv, RMS_norm, SiLU, conv] + shortcut
 * (wan_vae.py:198-202,206-224), so each conv's output is immediately normalised over channels
 * (wan_vae.py:43-58).  Cout in {96, 192} (one epilogue thread owns a whole pixel); x bf16
 * [T, H, W, Cin]; residual (or NULL) is added to the conv output first.  Writes
 *   out      [T, H, W, Cout] = conv (+ residual)            — or skipped when out == NULL
 *   norm_out [T, H, W, Cout] = [SiLU](RMS_norm(out) * gamma), same rounding points as
 *                              m4d_rmsnorm_silu_cl applied to `out`. */
int m4d_conv3x3_rmsnorm_cl(const void* x, int T, int H, int W, int Cin, const void* w_packed, int Cout,
                           const void* bias, int kt, void* out, const void* residual,
                           const void* gamma, void* norm_out, int do_silu, void* stream);

/* RMS_norm over channels (* sqrt(C) * gamma) [+ SiLU] per pixel, in place allowed
 * (wan_vae.py:43-58 + nn.SiLU at :198-202). */
int m4d_rmsnorm_silu_cl(const void* x, const void* gamma, void* out, long long pixels, int C,
                        int do_silu, void* stream);

/* nearest-exact 2x spatial upsample [T,H,W,C] -> [T,2H,2W,C] (wan_vae.py:61-67). */
int m4d_upsample2x_cl(const void* x, void* out, int T, int H, int W, int C, void* stream);

/* planar [C, P] -> channels-last [P, Cpad]; with div/add: z / inv_std + mean (wan_vae.py:682-686). */
int m4d_planar_to_cl(const void* x, void* out, long long P, int C, int Cpad, const float* div,
                     const float* add, void* stream);

/* channels-last [P, C] -> planar [C, P]; first n_affine channels get (v - sub) * mul
 * (wan_vae.py:540-545). */
int m4d_cl_to_planar(const void* x, void* out, long long P, int C, int n_affine, const float* sub,
                     const float* mul, void* stream);

/* 3x3 Conv2d (per frame, padding 1) over channels-last [T, H, W, Cin] -> [T, H, W, 128] whose epilogue
 * also produces the per-frame GroupNorm(32) statistics of its (bias / residual-added, bf16-rounded)
 * output — the input of the next Normalize in the adaptors' ResnetBlock chain
 * (trajectory_module.py:54-60,104-122) — so that the following m4d_groupnorm_apply_cl needs no
 * statistics pass.  partials_ws: T * ceil(H/16) * ceil(W/16) * 64 floats of scratch; stats:
 * T * M4D_GN_SLICES * 64 floats (sum, sum of squares per group and slice; deterministic). */
#define M4D_GN_SLICES 8
int m4d_conv3x3_gnstats_cl(const void* x, int T, int H, int W, int Cin, const void* w_packed, int Cout,
                           const void* bias, void* out, const void* residual, float* partials_ws,
                           float* stats, void* stream);

/* The apply pass of m4d_groupnorm_swish_cl with statistics given as `slices` partial
 * (sum, sum of squares) sets per frame: stats[f][slice][group][2]. */
int m4d_groupnorm_apply_cl(const void* x, const void* weight, const void* bias, void* out,
                           const float* stats, int slices, int F, int HW, int C, int groups, float eps,
                           void* stream);

/* GroupNorm(32, eps, affine) + x*sigmoid(x) on [F, HW, C] channels-last
 * (trajectory_module.py:54-60); stats_ws: 64*F floats of scratch. */
int m4d_groupnorm_swish_cl(const void* x, const void* weight, const void* bias, void* out,
                           float* stats_ws, int F, int HW, int C, int groups, float eps,
                           void* stream);

/* p bf16 [rows, N] = softmax(s fp32 [rows, N] * scale) — VAE AttentionBlock (wan_vae.py:257). */
int m4d_softmax_rows(const float* s, void* p, int rows, int N, long long lds, long long ldp,
                     float scale, void* stream);

/* out[c, r] = in[r, c] (bf16). */
int m4d_transpose_bf16(const void* in, void* out, int R, int C, long long ld_in, long long ld_out,
                       void* stream);

/* out bf16 = SiLU(x fp32), n % 4 == 0 — input of the Motion-Perception-Module projection
 * (wan_transformer4d.py:746-748,781). */
int m4d_silu_bf16(const float* x, void* out, long long n, void* stream);

/* ---- Motion-Perception-Module front end ---- */

/* rows[(f*H+y)*W+x][tap*C + c] = act(in[f][y+dy-1][x+dx-1][c]) (zero outside), tap = dy*3+dx:
 * the im2col operand that turns each Conv2d(768,768,3,padding=1) of `feature_adapter`
 * (wan_transformer4d.py:888-892, applied at :1148 to the OmniMAE tokens viewed as [B,14,14,768])
 * into one m4d_gemm_bf16 with the tap-major packed weight.  do_silu applies the nn.SiLU between
 * the two convolutions on load (bf16 result, like the reference's bf16 activation).  C % 8 == 0. */
int m4d_im2col3x3_cl(const void* in, void* rows, int F, int H, int W, int C, int do_silu, void* stream);

/* F.interpolate(size=(H,W), mode='bilinear', align_corners=False) of in [B,h,w,C] channels-last
 * (fp32 math, bf16 result) written to all T frames of out [B,T,H,W,C] — wan_transformer4d.py:
 * 1149-1151 (`.unsqueeze(2).repeat(1,1,latent_T,1,1).flatten(2).transpose(1,2)` gives exactly
 * this layout, [B, T*H*W, C]).  out_silu (or NULL) additionally receives bf16(SiLU(out)), the first
 * op of every block's SpatialGuidanceModule.spatial_guide (:746-749); out may be NULL. */
int m4d_bilinear_repeat_cl(const void* in, void* out, void* out_silu, int B, int h, int w, int C, int T,
                           int H, int W, void* stream);

/* ---- stage hand-off: point-cloud projection with a z-buffer ---- */

/* render_with_project, scripts/inference/infer.py:222-258 (project() MoRe4D/utils/project_utils.py:
 * 47-71; torch.unique + index_reduce_('amin') z-buffer :238-241; torch_scatter mean :246; pad,
 * [W,H,3]->[H,W,3] transpose, uint8 truncation and hole mask :247-256) as three passes.
 * points, colors: DEVICE fp32 [N, 3] (colours 0..255); world2cam: HOST fp32 [4,4] row-major =
 * extrinsic.inverse(); intrinsic: HOST fp32 [3,3] (normalised image coordinates).  Outputs:
 * image DEVICE uint8 [H, W, 3], mask DEVICE uint8 [H, W] (1 = no point landed there).
 * workspace: DEVICE, 16-byte aligned, >= m4d_project_points_workspace(N, H, W) bytes. */
long long m4d_project_points_workspace(long long N, int H, int W);
int m4d_project_points(const float* points, const float* colors, const float* world2cam,
                       const float* intrinsic, long long N, int H, int W, unsigned char* image,
                       unsigned char* mask, void* workspace, long long workspace_bytes, void* stream);

/* The same for the V frames of a camera trajectory in ONE launch sequence (render_trajectory,
 * scripts/inference/infer.py:398-444, calls render_with_project once per frame: launch-bound).
 * points / colors: view v starts at element offset v * *_view_stride (0 = shared by all views);
 * world2cam: HOST fp32 [V, 4, 4]; image [V, H, W, 3], mask [V, H, W].  Bit-identical to V calls of
 * m4d_project_points. */
long long m4d_project_views_workspace(long long N, int V, int H, int W);
int m4d_project_views(const float* points, long long points_view_stride, const float* colors,
                      long long colors_view_stride, const float* world2cam, const float* intrinsic,
                      long long N, int V, int H, int W, unsigned char* image, unsigned char* mask,
                      void* workspace, long long workspace_bytes, void* stream);

/* Forward 3D-Gaussian-splatting render of V views in one launch sequence — `gs_render` /
 * `render_cuda` (MoRe4D/utils/gaussian_splatting.py:13-43,201-281) + the third-party
 * GaussianRasterizer it calls once per frame (README.md:60; scripts/inference/infer.py:260-273,
 * 398-444).  Published algorithm of that extension: projection + EWA covariance with 0.3 px
 * dilation, 16x16 tiles, depth-sorted front-to-back alpha blending (1/255 and 1e-4 cut-offs).
 * means: DEVICE fp32 [V, N, 3] (view stride in elements, 0 = shared); colors: DEVICE fp32 [N, 3] in
 * 0..1 (view stride likewise; precomputed colours, use_sh=False); opacity: DEVICE fp32 [N]; cov3d:
 * DEVICE fp32 [6] = xx xy xz yy yz zz of the ONE covariance all gaussians share
 * (build_covariance(scale, rotation), :140-151); cams: DEVICE fp32 [V, 20], 16-byte aligned: rows of
 * the world->camera [R|t] (12), fx fy in pixels, tan(fov_x/2) tan(fov_y/2), then P00 P11 P22 P23 of
 * get_projection_matrix (:198-226).  image: DEVICE fp32 [V, 3, H, W]; image_u8 (or NULL): DEVICE
 * uint8 [V, H, W, 3] = trunc(image * 255) (infer.py:272-273).  workspace: DEVICE, 16-byte aligned,
 * >= m4d_gs_render_workspace(N, V, H, W, dup_capacity) bytes, dup_capacity = room for (tile,
 * gaussian) pairs; *dup_needed (HOST, may be NULL) receives the count the call needs, and
 * M4D_ERR_WORKSPACE is returned if it exceeds dup_capacity (nothing is rendered then).  The call
 * synchronises the stream once (to read that count), like the original extension. */
long long m4d_gs_render_workspace(long long N, int V, int H, int W, long long dup_capacity);
int m4d_gs_render(const float* means, long long means_view_stride, const float* colors,
                  long long colors_view_stride, const float* opacity, const float* cov3d, const float* cams,
                  long long N, int V, int H, int W, float bg0, float bg1, float bg2, float* image,
                  unsigned char* image_u8, void* workspace, long long workspace_bytes, long long dup_capacity,
                  long long* dup_needed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MORE4D_B200_H */
