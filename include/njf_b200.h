/* njf_b200.h -- C ABI of libnjf_b200.so: the B200-native (sm_100a) volumetric-rendering hot path
 * of neural-jacobian-field.
 *
 * The reference (sizhe-li/neural-jacobian-field) has NO native boundary: its plugin surface is
 * the Python `Model` nn.Module plus the DENSITY_DECODERS / ACTION_DECODERS registries
 * (project/neural_jacobian_field/models/decoder/__init__.py:11-44).  The host-side mirror
 * (`neural-jacobian-field_b200/njf_b200`) keeps that surface and calls THIS library for all
 * rendering arithmetic.  Every entry point cites the reference code it replaces (paths relative
 * to project/neural_jacobian_field/).
 *
 * Conventions: plain pointers and sizes only; all tensors fp32, contiguous, row-major; pointers
 * are DEVICE pointers unless a parameter says "host"; every call enqueues on `stream`
 * (a cudaStream_t passed as void*) and returns 0 on success, non-zero on failure with a message
 * in njf_last_error().  Nothing is allocated behind the caller's back: NjfField owns only the packed
 * weights it is created with; every pass writes into caller-provided outputs and a caller-provided
 * `workspace` (njf_workspace_bytes), so a frame can be captured into a CUDA graph and two streams can
 * render from one NjfField concurrently as long as each uses its own workspace.
 * There is no CPU fallback anywhere in this library.
 */
#ifndef NJF_B200_H_
#define NJF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NJF_MAX_LEVELS 4 /* proposal levels (the shipped configs use 1: rendering.num_proposal_samples=[256]) */

const char* njf_last_error(void);
int njf_version(void);

/* ---- field = packed weights of the decoder + proposal networks ------------------------------
 * Replaces: nn.Module parameter storage of ActionDecoderJacobian{MLP,Transformer}
 * (models/decoder/action_decoder_jacobian.py:261-416) and DensityDecoderMlp
 * (models/decoder/density_decoder.py:23-43); architecture constants are the shipped ones
 * (configurations/model/model_allegro.yaml, model_toy_arm.yaml): ResnetFC 5 blocks x 128,
 * combine_layer 3, 10 frequencies, geometry_feature_dim 15, transformer 64/64/8 heads/3 layers/64.
 * Anything else is rejected. */
typedef struct NjfField NjfField;

enum { NJF_HEAD_TRANSFORMER = 0, NJF_HEAD_MLP = 1 };

typedef struct NjfFieldDesc {
  int head;          /* NJF_HEAD_*  (ACTION_DECODERS key "jacobian_transformer" / "jacobian_mlp") */
  int action_dim;    /* A: transformer head A <= 8, MLP head A <= 10 */
  int n_proposal;    /* number of proposal networks, 1..NJF_MAX_LEVELS */
  int encoder_dim;   /* must be 512 (EncoderResnet.get_output_dim, models/encoder/encoder_resnet.py:88) */
  int sh_fp16_round; /* 1: round the SH-16 direction encoding through fp16 like tiny-cuda-nn does (the encoding is an
                        fp16 tensor-core operand of the colour head here, so it is rounded either way) */
  int sh_convention; /* NJF_SH_TCNN (default, what SHEncoding(implementation="tcnn") computes: signs of tiny-cuda-nn,
                        input re-mapped x*2-1) or NJF_SH_NERFSTUDIO_TORCH (nerfstudio's torch fallback
                        components_from_spherical_harmonics: its sign set, evaluated on the [0,1] input as passed) */
} NjfFieldDesc;
enum { NJF_SH_TCNN = 0, NJF_SH_NERFSTUDIO_TORCH = 1 };

typedef struct NjfTensor {
  const char* name;  /* reference state-dict key without the "model." prefix, e.g.
                        "decoder.density_head.lin_z.0.weight", "proposal_networks.0.density_head.lin_in.bias" */
  const float* data; /* HOST pointer, fp32, contiguous, the module's native shape */
  int64_t numel;
} NjfTensor;

/* Packs the named fp32 host tensors into device-resident tcgen05 operand images (fp16, K-major,
 * 128B-swizzled), folds the cross-attention key/value projections of the index embedding, and
 * builds the per-kernel step programs.  Synchronous. */
int njf_field_create(const NjfFieldDesc* desc, const NjfTensor* tensors, int n_tensors, NjfField** out);
void njf_field_destroy(NjfField* f);
/* Partial re-pack after an optimiser step of the action phase (only `decoder.jacobian_*` changed,
 * models/model_wrapper.py:75-85): `tensors` = the cross-attention head's tensors (jacobian_query_mlp, jacobian_index_embedding,
 * jacobian_attn_decoder.*, jacobian_head).  Re-packs the head blob, the query step image and the 64 hoisted query rows in
 * place (synchronises `stream` first; no allocation).  Cross-attention head only. */
int njf_field_update_head(NjfField* f, const NjfTensor* tensors, int n_tensors, void* stream);

/* ---- hoisted feature maps ---------------------------------------------------------------------
 * Replaces the three `lin_z[k](z)` Linear(512->128) layers of every ResnetFC
 * (model_components/resnet_fc.py:139-143) and the 512-column part of jacobian_query_mlp
 * (action_decoder_jacobian.py:423-430): bilinear interpolation is linear, so these layers are
 * applied ONCE per image to the encoder output (NCHW fp32, models/encoder/encoder_resnet.py:53-86)
 * and the render kernels gather the transformed channels.  Output: fp16 pixel-major maps. */
size_t njf_hoisted_bytes(const NjfField* f, int B, int Hf, int Wf);
int njf_hoist_features(const NjfField* f, const float* feat_nchw, int B, int Hf, int Wf, void* maps_out,
                       void* stream);
/* the same for views [view0, view0 + B_local) of a B_total-view map buffer (njf_hoisted_bytes(f, B_total, ...)):
 * a rank of a ray-sharded multi-view call encodes and hoists only the views its ray range touches */
int njf_hoist_features_views(const NjfField* f, const float* feat_nchw, int B_local, int view0, int B_total, int Hf,
                             int Wf, void* maps_out, void* stream);
/* The same from a channels-last HALF-PRECISION encoder output: feat [B_local][Hf*Wf][512] fp16 (NHWC,
 * 16-byte aligned) -- already the K-major operand layout, so the kernel copies 16-byte chunks instead of transposing
 * and converting, and reads half the bytes. */
int njf_hoist_features_nhwc16(const NjfField* f, const void* feat_nhwc_f16, int B_local, int view0, int B_total, int Hf,
                              int Wf, void* maps_out, void* stream);

/* ---- cameras ---------------------------------------------------------------------------------- */
typedef struct NjfCameras {
  const float* ctxt_w2c;  /* [B][16] inverse of CameraInput.ctxt_extrinsics (rendering/geometry.py:59-65) */
  const float* ctxt_k;    /* [B][9]  normalised intrinsics (pixel_aligned_features.py:21) */
  const float* trgt_w2c;  /* [B][16] inverse of CameraInput.trgt_extrinsics (geometry.py:206-215) */
  const float* trgt_k_px; /* [B][9]  pixel-unit intrinsics (models/model.py:305-312) */
  /* optional HOST copies of ctxt_w2c / ctxt_k (NULL if unknown): when present together with
     NjfRenderArgs.h_z_near / h_z_far and B <= 16, the per-view constants ride in the kernel
     parameter block (constant bank) and row set-up issues no global loads for them */
  const float* h_ctxt_w2c;
  const float* h_ctxt_k;
} NjfCameras;

/* ---- Model.forward / encode_image ------------------------------------------------------------
 * Replaces models/model.py:316-396 (forward, eval mode) and :458-495 (encode_image):
 * ProposalNetworkSampler.generate_ray_samples (rendering/ray_samplers.py:497-552), the decoder
 * forward (action_decoder_jacobian.py:147-215), RaySamples.get_weights (:77-101) and the
 * render_* reductions (model.py:257-314).  Any output pointer may be NULL. */
typedef struct NjfRenderArgs {
  int B, R;                       /* views, rays per view */
  int n_levels;                   /* proposal levels */
  int s_prop[NJF_MAX_LEVELS];     /* RenderingCfg.num_proposal_samples */
  int s_nerf;                     /* RenderingCfg.num_nerf_samples */
  const float* origins;           /* [B][R][3] */
  const float* dirs;              /* [B][R][3] */
  const float* z_near;            /* [B] */
  const float* z_far;             /* [B] */
  const float* h_z_near;          /* optional HOST copy of z_near (see NjfCameras.h_ctxt_w2c) */
  const float* h_z_far;           /* optional HOST copy of z_far */
  const float* action;            /* [B][A] */
  const float* bins0;             /* level-0 spacing bins: [s_prop[0]+1] shared (stride 0) or per ray */
  int bins0_stride;               /* 0 or s_prop[0]+1 (train-mode stratified jitter comes in this way) */
  const float* u[NJF_MAX_LEVELS]; /* PDF sample positions per resampling level: [n+1], n = next level's count */
  int u_stride[NJF_MAX_LEVELS];   /* 0 (shared, eval mode) or n+1 (per ray, train mode) */
  float anneal;                   /* ProposalNetworkSampler._anneal (1.0 at inference) */
  int sum_vec_width;              /* 0: exact sum; 8/16: reproduce ATen's vectorised CPU fp32 sum order */
  const void* maps;               /* njf_hoist_features output */
  int Hf, Wf;
  /* per-ray outputs */
  float* rgb;                     /* [B][R][3]   render_rgb */
  float* depth;                   /* [B][R][1]   render_depth (clipped to the call-global [min,max] of steps) */
  float* flow;                    /* [B][R][2]   render_optical_flow */
  float* jbar;                    /* [B][R][3A]  render_action_features */
  float* p;                       /* [B][R][3]   sum w x */
  float* pw;                      /* [B][R][3]   sum w (x + J u) */
  /* per-sample outputs (compute_vis_features / encode_image) */
  float* steps;                   /* [B][R][S] */
  float* weights;                 /* [B][R][S] */
  float* sigma;                   /* [B][R][S] */
  float* jac;                     /* [B][R][S][3A] */
  float* positions;               /* [B][R][S][3] */
  float* rgb_samples;             /* [B][R][S][3] */
  /* sampler intermediates (ModelTrainingOutput / tests) */
  float* prop_weights[NJF_MAX_LEVELS]; /* [B][R][s_prop[l]] */
  float* level_bins[NJF_MAX_LEVELS];   /* REQUIRED workspace: [B][R][n_l+1], bins produced by level l's PDF step */
  int32_t* level_inds[NJF_MAX_LEVELS]; /* [B][R][n_l+1] searchsorted results */
  float* minmax;                  /* REQUIRED workspace: 2 floats; after njf_field_pass = (min, max) of steps over the call
                                     (all-reduce it across ranks before njf_finish_pass when one call is ray-sharded) */
  /* caller-owned scratch ("caller allocates outputs and a workspace"): holds delta*sigma between
     proposal_kernel and pdf_kernel (when prop_weights[l] is NULL) and the fp16 query-embedding hand-over between
     field_kernel and xf_kernel (cross-attention head).  Any size >= njf_workspace_min_bytes works: a pass whose
     intermediates do not fit runs as several launch groups over ray ranges (bit-identical results);
     njf_workspace_bytes returns the size at which no pass is split more than needed to keep the hand-over
     <= 256 MiB.  16-byte aligned. */
  void* workspace;
  size_t workspace_bytes;
  float* packed;                  /* optional [B][R][12+3A]: rgb3 | depth1 | flow2 | jbar3A | p3 | pw3 per ray, written by
                                     njf_finish_pass (one buffer for the multi-GPU gather) */
  /* ray-sharded call ("flatten (view, ray) and split contiguously by rank"): when n_rays > 0 this call
     renders only rays [ray_offset, ray_offset + n_rays) of the flattened B x R index space.  origins / dirs /
     bins0 / u (per-ray strides) and every per-ray / per-sample output then hold n_rays entries; the per-view
     inputs (cameras, z_near, z_far, action, maps) stay arrays over all B views (maps of views the range does not
     touch are never read).  n_rays == 0: the whole B x R call. */
  int ray_offset;
  int n_rays;
} NjfRenderArgs;

/* workspace sizing for a (B, R, sample-count) render with this field */
size_t njf_workspace_bytes(const NjfField* f, int B, int R, int n_levels, const int* s_prop, int s_nerf);
size_t njf_workspace_min_bytes(const NjfField* f, int n_levels, const int* s_prop, int s_nerf);

int njf_render_forward(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* args, void* stream);

/* Stage entry points (the same kernels njf_render_forward chains; exposed for stage-wise parity). */
int njf_proposal_pass(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* args, int level,
                      const float* bins_in, int bins_in_stride, void* stream);
int njf_field_pass(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* args, const float* bins,
                   int bins_stride, void* stream);
int njf_finish_pass(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* args, void* stream);

/* ---- Model.compute_density (models/model.py:416-456) and the decoder plugin surface
 * ActionDecoderJacobian.forward / encode_image / compute_density (action_decoder_jacobian.py:92-249): density head,
 * Jacobian head and colour head at explicit world-space points (no rays).  points [B][N][3]; dirs [B][N][3] unit
 * view directions (NULL unless rgb is wanted); sigma [B][N], geo [B][N][15], jac [B][N][3A], rgb [B][N][3]; any
 * output may be NULL. */
int njf_query_points(const NjfField* f, const float* ctxt_w2c, const float* ctxt_k, const void* maps, int Hf,
                     int Wf, const float* points, const float* dirs, int B, int N, float* sigma, float* geo,
                     float* jac, float* rgb, void* workspace, size_t workspace_bytes, void* stream);
/* DensityDecoderMlp.get_density (models/decoder/density_decoder.py:45-71) of proposal network `level` at explicit
 * points: sigma [B][N] */
int njf_query_proposal_density(const NjfField* f, int level, const float* ctxt_w2c, const float* ctxt_k,
                               const void* maps, int Hf, int Wf, const float* points, int B, int N, float* sigma,
                               void* stream);
/* workspace for njf_query_points (cross-attention head hand-over; 0 for the MLP head) */
size_t njf_query_workspace_bytes(const NjfField* f, int B, int N);
/* by-products returned by the reference's DensityHeadOutput: positional encoding (63) of the
 * context-camera point and the bilinear gather of the RAW encoder features (NCHW fp32, C channels);
 * either output may be NULL.  (pixel_aligned_features.py:11-35, action_decoder_jacobian.py:97-104) */
int njf_point_features(const float* feat_nchw, const float* ctxt_w2c, const float* ctxt_k, const float* points,
                       int B, int N, int C, int Hf, int Wf, float* xyz_features, float* pixel_aligned_features,
                       void* stream);

/* ---- inverse of n 4x4 poses (torch.inverse in rendering/geometry.py:59-65 transform_world2cam and :206-215):
 * Gauss-Jordan with partial pivoting in fp64, rounded to fp32 -- on the device, so that a frame needs no host
 * LAPACK call and no host->device copy of the inverted matrices.  in/out [n][16]. */
int njf_invert_poses(const float* c2w, float* w2c, int n, void* stream);

/* ---- ray bundle of a view (rendering/geometry.py:117-134 get_pixel_coordinates, :170-203 get_world_rays_with_z,
 * models/model.py:215-226): k_norm [B][9] normalised intrinsics, c2w [B][16]; coords_xy [B][R][2] normalised pixel
 * coordinates, or NULL to generate the H x W grid of pixel centres in the reference's order (row-major, x fastest;
 * R must equal H*W).  origins / dirs [B][R][3]; z [B][R] (camera-space z of the unit direction) may be NULL. */
int njf_make_rays(const float* k_norm, const float* c2w, const float* coords_xy, int B, int R, int H, int W,
                  float* origins, float* dirs, float* z, void* stream);

/* ---- PDFSampler alone (rendering/ray_samplers.py:351-451, eval/train u supplied by the caller); s_in <= 4096.
 * With sum_vec_width = 8 the fp32 row sum reproduces ATen's CPU order bit for bit for rows of up to 512 samples (the
 * render path's limit). */
int njf_pdf_sample(const float* weights, const float* bins_in, int bins_in_stride, const float* u, int u_stride,
                   int n_rays, int s_in, int n_out, float anneal, int sum_vec_width, float* bins_out,
                   int32_t* inds_out, void* stream);

/* ---- RaySamples.get_weights alone (rendering/ray_samplers.py:77-101) */
int njf_transmittance_weights(const float* deltas, const float* sigma, int n_rays, int s, float* weights_out,
                              void* stream);

/* ---- Model.infer_optical_flow (models/model.py:497-525) on the collapsed encoding:
 * jbar [N][3A], p [N][3], action [B][A] (rays_per_view rays per action row) -> flow [N][2], pw [N][3] */
int njf_flow_from_encoding(const float* jbar, const float* p, const float* action, const float* trgt_w2c,
                           const float* trgt_k_px, int n_rays, int rays_per_view, int action_dim, float* flow,
                           float* pw, void* stream);

/* ---- action-phase training: backward of the cross-attention Jacobian head.
 * models/model_wrapper.py:75-85 freezes everything but the Jacobian head and :148-163 puts an MSE on the optical
 * flow, so the gradient path is  flow -> pw = p + Jbar^T u -> Jbar = sum_s w_s J_s -> J_s = head(q0_s) -> W_q, b_q.
 *
 * njf_flow_backward: backward of njf_finish_pass / njf_flow_from_encoding (models/model.py:288-314, 497-525):
 *   g_flow [N][2] (+ optional g_pw_in [N][3]) -> g_jbar [N][3A] (nullable), g_action [B][A] (nullable, zeroed here).
 * njf_xf_backward: given the hand-over region a train-mode njf_field_pass left in ITS workspace (single launch
 *   group: workspace >= all tiles; fp16 query embedding + sample weights per 128-row tile) and g_jbar, recomputes
 *   the three attention / feed-forward layers in fp32 and back-propagates.  `folded` / `g_folded`
 *   (njf_xf_folded_floats() floats, fp32, natural-base softmax) hold, per layer,
 *   [M1 64x64 | m1b | M2 | b_out | W1 | w1b | W2 | b2] and then [W_head 64x64 (rows >= 3A zero) | b_head 64], i.e.
 *   the matrices field.cu folds from the state dict (transformer.py:14-21, 63-82): rows of M1 / columns of M2 are
 *   indexed h*8 + a.  g_folded is ACCUMULATED into.  g_q0 [n_tiles*128][64] receives d loss / d q0 per row.
 * njf_query_backward: q0 = W_q [enc63 | feat512] + b_q (action_decoder_jacobian.py:423-430): g_bq [64],
 *   g_wq_enc [64][64] (columns 0..62 = d W_q[:, :63]), both accumulated, and g_map [B][Hf*Wf][64] (accumulated;
 *   the adjoint of the bilinear gather), from which d W_q[:, 63:] = sum_pixels g_map^T . features. */
int njf_xf_folded_floats(void);
size_t njf_xf_backward_workspace_bytes(int n_tiles);
int njf_xf_backward(const float* folded, int action_dim, const void* handover, int n_tiles, int n_rays, int s_nerf,
                    const float* g_jbar, float* g_folded, float* g_q0, void* workspace, size_t workspace_bytes,
                    void* stream);
int njf_query_backward(const NjfCameras* cams, const NjfRenderArgs* args, const float* final_bins, int bins_stride,
                       const float* g_q0, int n_tiles, float* g_wq_enc, float* g_bq, float* g_map, void* stream);
int njf_flow_backward(const float* g_flow, const float* g_pw_in, const float* jbar, const float* p, const float* action,
                      const float* trgt_w2c, const float* trgt_k_px, int n_rays, int rays_per_view, int action_dim,
                      float* g_jbar, float* g_action, void* stream);

/* ---- training of the ResnetFC trunks: perception phase (models/model_wrapper.py:116-146: density
 * head, colour head, proposal networks, the encoder through the feature map) and the MLP Jacobian head of the action
 * phase.  The fused render kernels keep no activations, so a training step evaluates the trunks layer by layer in
 * fp32 with the activations in caller-owned HBM tensors; torch.autograd chains these entry points
 * (njf_b200/train_trunk.py) the way it chains nn.Linear in the reference (model_components/resnet_fc.py:70-79,
 * 130-154).  lin_z is hoisted onto the feature map exactly as at inference: lin_z(bilinear(f)) = bilinear(lin_z(f)),
 * the per-pixel maps being one plain GEMM over the NHWC encoder output.  Every matrix is row-major, contiguous, fp32;
 * inner sizes are multiples of 4 in [4,128] (the caller zero-pads 63 -> 64, 31 -> 32, 3 -> 4, ...).
 *
 * njf_train_sample_setup: world points [B][N][3] -> context-camera point -> NeRFEncoding enc [B*N][64] (63 columns in
 *   nerfstudio's order + one zero) and the four bilinear taps of F.grid_sample(align_corners=True, border)
 *   (pixel_aligned_features.py:11-35): tap_pix [B*N][4] pixel indices view*Hf*Wf + y*Wf + x, tap_w [B*N][4].
 * njf_train_gather / njf_train_scatter: a window [ch0, ch0 + CW) of the CH map channels (one lin_z layer = 128 of 384):
 *   out[m][c] = sum_t tap_w[m][t] map[tap_pix[m][t]][ch0 + c], out [M][CW], and its adjoint
 *   dmap[tap_pix[m][t]][ch0 + c] += tap_w[m][t] g[m][c], g [M][CW] (dmap [pixels][CH] is accumulated into); CW a multiple of 128.
 * njf_train_linear: c[M][n_out] = mask(act(a[M][k_red]) . Wm + bias) + residual, act = ReLU when relu_in, mask zeroes
 *   entries whose mask_src[M][n_out] <= 0; bias / residual / mask_src may be NULL.
 *   trans_w = 1: w is a Linear weight [n_out][k_red] (forward, y = x W^T + b);
 *   trans_w = 0: w is the same weight seen as [k_red][n_out] (input gradient g_x = g_y W, masked by ReLU' of the saved x).
 * njf_train_linear_wgrad: gw[N][K] += gy[M][N]^T . act(x[M][K]), gb[N] += column sums of gy (gb may be NULL).
 * tensor_cores = 0: fp32 SIMT kernels (exact to fp32 round-off); tensor_cores = 1: tcgen05 kind::tf32 with fp32
 *   accumulation in tensor memory -- what the reference's nn.Linear layers run under
 *   torch.set_float32_matmul_precision("high") (train.py:64-65); the Python layer follows that torch switch.
 * njf_train_sh16: SH degree 4 of unit directions dirs [M][3] -> out [M][16] (action_decoder_jacobian.py:24-30, 284),
 *   optionally rounded through fp16 like tiny-cuda-nn's output. */
int njf_train_sample_setup(const float* ctxt_w2c, const float* ctxt_k, const float* points, int B, int N, int Hf, int Wf,
                           float* enc, int* tap_pix, float* tap_w, void* stream);
int njf_train_gather(const float* map, const int* tap_pix, const float* tap_w, int M, int CH, int ch0, int CW, float* out,
                     void* stream);
int njf_train_scatter(const float* g, const int* tap_pix, const float* tap_w, int M, int CH, int ch0, int CW, float* dmap,
                      void* stream);
int njf_train_linear(const float* a, const float* w, const float* bias, const float* residual, const float* mask_src,
                     float* c, int M, int n_out, int k_red, int trans_w, int relu_in, int tensor_cores, void* stream);
int njf_train_linear_wgrad(const float* gy, const float* x, int M, int N, int K, int relu_in, float* gw, float* gb,
                           int tensor_cores, void* stream);
int njf_train_sh16(const float* dirs, int M, int sh_convention, int fp16_round, float* out, void* stream);

/* ---- inverse dynamics on the collapsed encoding (the Adam loop of notebooks/real_world/2_inverse_dynamics.ipynb
 * over Model.infer_optical_flow, models/model.py:497-525): Gauss-Newton normal equations of
 *   min_u sum_i w_i | flow_i(u) - target_i |^2 ,  flow_i(u) = proj(p_i + Jbar_i^T u) - proj(p_i)
 * per view: H [B][A][A] = sum w G^T G, g [B][A] = sum w G^T r, loss [B] = sum w |r|^2 (fp64), G = d flow / d u.
 * target_flow [N][2] pixels; ray_weight [N] or NULL; workspace: njf_flow_gn_workspace_doubles(B) doubles. */
int njf_flow_gn_terms(const float* jbar, const float* p, const float* action, const float* trgt_w2c,
                      const float* trgt_k_px, const float* target_flow, const float* ray_weight, int n_rays,
                      int rays_per_view, int action_dim, double* workspace, double* H, double* g, double* loss,
                      void* stream);
int njf_flow_gn_workspace_doubles(int n_views);

/* ---- self-test of the tcgen05 layer-chain machinery (tests only; weights/bias are HOST pointers) */
int njf_selftest_chain(const float* w0, const float* w1, const float* w2, const float* w3, const float* bias,
                       const float* a_in, const float* tz_in, float* x_out, float* y_out, int ntiles, int grid,
                       void* stream);

/* ---- measurement aid (bench.py): with enable != 0 every njf_field_pass brackets its kernels with CUDA events on
 * the launching stream; a later call returns the accumulated milliseconds of field_kernel and of xf_kernel (the
 * cross-attention head; 0 for the MLP head) since the previous call and resets the counters.  Either pointer may be NULL. */
int njf_debug_field_timing(int enable, float* field_kernel_ms, float* xf_kernel_ms);
/* number of kernels the library has launched since the last reset (bench.py's gpu_launches) */
long long njf_debug_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* NJF_B200_H_ */
