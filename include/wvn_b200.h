/* wvn-b200 — C ABI of libwvn_b200.so
 *
 * B200-native (sm_100a) implementation of the Wild Visual Navigation per-frame hot path:
 * DINO ViT forward -> STEGO segmentation head / per-segment feature gather -> traversability
 * MLP (per-pixel inference with confidence, and the online train step).
 *
 * The reference (leggedrobotics/wild_visual_navigation @ 8b9caf9) is pure Python and has no
 * FFI for this path: its boundary is the Python class surface
 * (FeatureExtractor / DinoInterface / StegoInterface / SegmentExtractor / SimpleMLP /
 * TraversabilityLoss / ConfidenceGenerator / TraversabilityEstimator).  The entry points
 * below are what a ctypes binding inside those classes calls; each cites the reference
 * code it replaces (paths relative to the reference repository root).  INTEGRATION.md shows
 * the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless named host_*; tensors are dense, row-major;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *  - functions return 0 on success, a negative wvn_status otherwise; wvn_last_error() gives
 *    the message for the calling thread;
 *  - no ownership transfer; handles own their weights and workspaces (allocated at create,
 *    nothing is allocated on the per-frame path);
 *  - a handle must not be used from two host threads at once (the reference serialises
 *    frames through its Scheduler and training through `_learning_lock`).
 */
#ifndef WVN_B200_H_
#define WVN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  WVN_STATUS_OK = 0,
  WVN_STATUS_INVALID = -1,   /* bad argument / unsupported shape */
  WVN_STATUS_CUDA = -2,      /* CUDA error, see wvn_last_error() */
  WVN_STATUS_NO_DEVICE = -3, /* no sm_100 device visible */
  WVN_STATUS_STATE = -4      /* handle not ready (e.g. a weight was never set) */
} wvn_status;

const char* wvn_last_error(void);
/* 0 if device 0..n has compute capability 10.x, WVN_STATUS_NO_DEVICE otherwise. */
int wvn_check_device(void);
int wvn_version(void);
/* Number of kernel launches this library has issued in this process (bench.py's gpu_launches). */
long long wvn_launch_count(void);
/* Optional CUDA-event timing of the dominant kernels inside real steps (bench.py roofline):
 * category 0 = fused attention, 1 = tcgen05 GEMMs; category_mask bit c enables category c (0 = off, 1 = the
 * roofline kernel only — 24 event pairs per step —, 3 = attention + the ~280 GEMM launches per step).
 * collect() synchronises and writes the summed milliseconds / launch counts per category into HOST arrays of
 * length 2 and clears the records. */
void wvn_profile_enable(int category_mask);
int wvn_profile_collect(float* host_ms, long long* host_launches);

/* ------------------------------------------------------------------------------------------
 * Primitive: bf16 tensor-core GEMM  C = A[M,K] * W[N,K]^T  (tcgen05, fp32 accumulate)
 * Replaces every nn.Linear / 1x1-conv call of the path (cuBLAS in the reference).
 * out_kind: 0 = bf16 out[M,ldo] = act(acc + bias); 1 = fp32 out = acc + bias;
 *           2 = fp32 out += acc + bias (residual).   act: 0 none, 1 ReLU, 2 GELU(erf).
 * K % 64 == 0, N % block_n == 0, block_n in {0 (auto), 64, 128, 192, 224, 256}.
 * ---------------------------------------------------------------------------------------- */
int wvn_gemm_bf16(const void* a_bf16, long long lda, const void* w_bf16, const float* bias, void* out, long long ldo,
                  int m, int n, int k, int out_kind, int act, int block_n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Primitive: fused multi-head attention, head dim 64, non-causal.
 * Replaces `softmax(q @ k.transpose(-2,-1) * scale) @ v` of the DINO ViT block
 * ([EXTERNAL] stego/backbones/dino/vision_transformer.py Attention.forward).
 * q,k: [batch*heads, npad, 64] bf16; vt: [batch*heads, 64, npad] bf16;
 * out: [batch, npad, heads*64] bf16.  Keys >= n_valid are masked.  npad % 128 == 0.
 * ---------------------------------------------------------------------------------------- */
int wvn_attention_bf16(const void* q, const void* k, const void* vt, void* out, int batch, int heads, int npad,
                       int n_valid, float scale, void* stream);

/* LayerNorm over fp32 rows -> bf16 (and/or fp32) rows; dim in {384, 768}. */
int wvn_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, long long rows, int dim,
                  float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * ViT backbone handle — replaces DinoInterface.__init__/inference's backbone
 * (wild_visual_navigation/feature_extractor/dino_interface.py:16-59, :70-92 and the
 *  [EXTERNAL] stego.backbones.backbone.get_backbone -> DINO VisionTransformer).
 * ---------------------------------------------------------------------------------------- */
typedef struct wvn_vit wvn_vit_t;

typedef struct {
  int image_size;   /* side of the (square) network input after resize + center crop, e.g. 448 */
  int patch_size;   /* 8 or 16 */
  int dim;          /* 384 (vit_small) or 768 (vit_base) */
  int depth;        /* 12 */
  int heads;        /* dim / 64 */
  int mlp_dim;      /* 4 * dim */
  int max_batch;    /* largest batch a single forward call may carry */
  int chunk;        /* frames processed together (activations stay L2-resident); 0 = default */
  float ln_eps;     /* 1e-6 */
  int head_out;     /* columns of the STEGO head output matrix (multiple of 64), 0 = no head */
} wvn_vit_config;

int wvn_vit_create(const wvn_vit_config* cfg, wvn_vit_t** out);
void wvn_vit_destroy(wvn_vit_t* h);

/* Weights use the DINO state-dict names (fp32, host or device memory):
 *   cls_token [dim], pos_embed [(1+P), dim] (already interpolated to this token grid),
 *   patch_embed.proj.weight [dim, 3*p*p], patch_embed.proj.bias [dim],
 *   blocks.{i}.norm1.{weight,bias}, blocks.{i}.attn.qkv.{weight [3dim,dim],bias},
 *   blocks.{i}.attn.proj.{weight,bias}, blocks.{i}.norm2.{weight,bias},
 *   blocks.{i}.mlp.fc1.{weight [mlp,dim],bias}, blocks.{i}.mlp.fc2.{weight [dim,mlp],bias},
 *   norm.{weight,bias}
 * and for the STEGO head (see wvn_vit_stego_head):
 *   stego.head_a.weight [head_out, dim], stego.head_a.bias [head_out],
 *   stego.hidden.weight [dim, dim], stego.hidden.bias [dim], stego.head_b.weight [head_out, dim]. */
int wvn_vit_set_weight(wvn_vit_t* h, const char* name, const float* data, long long numel);

/* img: [batch, 3, in_h, in_w] fp32 in [0,1].  Applies the interface's transform
 * (Resize(image_size, NEAREST) + CenterCrop(image_size) + ImageNet Normalize,
 *  dino_interface.py:52-59) inside the patch loader, runs the backbone, and writes the final
 * LayerNorm output of the patch tokens (CLS dropped): tokens_out [batch, P, dim] fp32.
 * resized_h/resized_w: size after the virtual NEAREST resize (== in_h/in_w when no resize). */
int wvn_vit_forward(wvn_vit_t* h, const float* img, int batch, int in_h, int in_w, int resized_h, int resized_w,
                    float* tokens_out, void* stream);

/* Flip test-time augmentation of STEGO ([EXTERNAL] Stego.get_code, called at stego_interface.py:91): the `batch`
 * frames are run twice, the second time on the horizontal flip of the TRANSFORMED (resized + cropped + normalised)
 * image; tokens_out [2*batch, P, dim] (second half = flipped pass) and the activation layout hold 2*batch frames,
 * so 2*batch <= max_batch. */
int wvn_vit_forward_tta(wvn_vit_t* h, const float* img, int batch, int in_h, int in_w, int resized_h, int resized_w,
                        float* tokens_out, void* stream);

/* Same as wvn_vit_forward, from the camera frame itself: img_hwc [batch, in_h, in_w, 3] uint8 RGB.  Folds into the patch loader the
 * two steps that precede the interface in the reference's node:
 *   ros_image_to_torch  — torchvision ToTensor: HWC uint8 -> CHW float / 255   (ros_converter.py:113-126)
 *   ImageProjector.resize_image — Resize(h, NEAREST) + CenterCrop(h)            (image_projector.py:55-59,199-200)
 * followed by the interface's own transform as in wvn_vit_forward; with network_input_image_height == image_size
 * (the shipped configuration) the two NEAREST resizes are one, expressed by resized_h / resized_w. */
int wvn_vit_forward_u8(wvn_vit_t* h, const unsigned char* img_hwc, int batch, int in_h, int in_w, int resized_h,
                       int resized_w, float* tokens_out, void* stream);

/* STEGO segmentation head on the tokens of the last forward
 * (stego_interface.py:91-100; [EXTERNAL] stego.stego.Stego.forward / cluster & linear probes):
 *   out[row, :] = head_a(t) + head_b(relu(hidden(t)))           out: [batch*npad, head_out] fp32
 * where the caller stacks code / cluster-probe / linear-probe rows into head_a / head_b.
 * Row b*npad + 1 + p holds patch p of frame b (row b*npad is the CLS token). */
int wvn_vit_stego_head(wvn_vit_t* h, int batch, float* out, void* stream);
int wvn_vit_npad(const wvn_vit_t* h);

/* ------------------------------------------------------------------------------------------
 * Token grid -> image resolution
 * ---------------------------------------------------------------------------------------- */
/* F.interpolate(features, (out,out), "bilinear", align_corners=True) (dino_interface.py:87-90).
 * tokens: [batch, gh*gw, dim] fp32 -> out: [batch, dim, out_h, out_w] fp32. */
int wvn_upsample_dense(const float* tokens, float* out, int batch, int dim, int gh, int gw, int out_h, int out_w,
                       void* stream);
/* bilinear (align_corners=False) upsampling of per-patch logits + argmax -> int64 segment ids
 * (STEGO postprocess + stego_interface.py:108-109).  logits: [batch*npad, ld] fp32.  Two logit
 * column ranges (cluster probe -> seg, linear probe -> seg_b; seg_b may be NULL) share one pass. */
int wvn_logits_argmax(const float* logits, long long ld, int col0, int classes, int col0_b, int classes_b, int batch,
                      int npad, int gh, int gw, int out_h, int out_w, long long* seg, long long* seg_b, void* stream);

/* Per-image k-means of the STEGO code (run_clustering=True — the reference's default for stego segmentation,
 * feature_extractor.py:47-53; [EXTERNAL] Stego.postprocess(image_clustering=True), stego_interface.py:91-100).
 * rows: the head-output matrix [batch*npad, ld] fp32 (row 0 of every frame = CLS), code in columns
 * [code_col, code_col+code_dim).  Lloyd iterations (Euclidean, `iters` fixed, evenly spaced deterministic init) over the
 * `patches` code rows of each frame; the per-patch nearest-centroid scores <x,c_k> - |c_k|^2/2 are written to columns
 * [logit_col, logit_col+k) so that wvn_logits_argmax yields the per-pixel labels of the upsampled code.
 * centroids_out: optional [batch, k, code_dim]; workspace: wvn_stego_kmeans_workspace_bytes(batch, k, code_dim) bytes of device
 * memory (partial sums of the 8 CTAs that share a frame; no initialisation needed).  One launch for all frames and iterations. */
/* STEGO's flip test-time augmentation (Stego.get_code: the code of the image and of its horizontal flip are averaged):
 * head [2*batch*npad, ld] holds the head output of the straight pass (frames [0, batch)) and of the pass over the flipped
 * transformed images (frames [batch, 2*batch), see wvn_vit_forward_tta).  In place: rows of the straight pass become
 * 0.5 * (own + mirrored row of the flipped pass); CLS / padding rows are zeroed. */
int wvn_flip_average(float* head, int batch, int npad, int grid, long long ld, void* stream);
size_t wvn_stego_kmeans_workspace_bytes(int batch, int k, int code_dim);
int wvn_stego_kmeans(float* rows, long long ld, int batch, int npad, int patches, int code_col, int code_dim, int logit_col,
                     int k, int iters, float* centroids_out, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Segment reductions — replace SegmentExtractor.adjacency_list / .centers
 * (feature_extractor/segment_extractor.py:40-92), FeatureExtractor.sparsify_features
 * (feature_extractor.py:389-396) and the relabel loop of segment_stego (:245-246).
 * ---------------------------------------------------------------------------------------- */
size_t wvn_segment_workspace_bytes(int batch, int smax, int gh, int gw);
/* seg: [batch, h, w] int64 with ids in [0, smax) (others ignored).
 * feat: [batch, smax, dim] fp32 or NULL (needs tokens [batch, gh*gw, dim] fp32);
 * centers: [batch, smax, 2] fp32 (x=col, y=row) or NULL;
 * edges: [batch, max_edges, 2] int64 (le, ri) sorted like torch.unique of the reference's keys,
 * n_edges: [batch] int32; both NULL to skip. */
int wvn_segment_reduce(const long long* seg, int batch, int h, int w, int smax, const float* tokens, int gh, int gw,
                       int dim, float* feat, float* centers, long long* edges, int* n_edges, int max_edges,
                       void* workspace, void* stream);
/* In-place relabel of each frame to 0..S-1; scratch: [batch*num_labels] int32; counts: [batch] int32. */
int wvn_segment_relabel(long long* seg, int batch, long long pix_per_frame, int num_labels, int* scratch, int* counts,
                        void* stream);

/* Supervision label pooling — replaces MissionNode.update_supervision_signal (traversability_estimator/nodes.py:400-440):
 *   signal = supervision_mask.nanmean(0);  per segment s: mean of signal over the segment's non-NaN pixels,
 *   nan_to_num(0);  valid = signal_mean > 0
 * without the reference's (H, W, S) one-hot expansion.  seg: [batch, h, w] int64 (ids outside [0, smax) ignored);
 * mask: [batch, channels, h, w] fp32 with NaN = unlabelled; y: [batch, smax] fp32; y_valid: [batch, smax] uint8;
 * count_ws: [batch, smax] fp32 scratch. */
int wvn_supervision_pool(const long long* seg, const float* mask, int batch, int channels, int h, int w, int smax,
                         float* y, unsigned char* y_valid, float* count_ws, void* stream);

/* SLIC superpixels — replaces fast_slic's Slic(num_components, compactness).iterate(np.uint8(img * 255)) behind
 * FeatureExtractor(segmentation_type="slic") (feature_extractor/feature_extractor.py:88-95, 221-225).  fast-slic is an
 * un-vendored C++ dependency; the algorithm is restated in all-integer form (oracle/slic.py is the definition; the labels
 * of this kernel are bit-identical to it).  img: [batch,3,h,w] fp32 in [0,1]; labels: [batch,h,w] int64 cluster ids in
 * [0, nx*ny) (empty clusters are possible: compact with wvn_segment_relabel); lut_*: device copies of the host tables
 * wvn_slic_tables() fills (256 / 9 / 4096 ints); workspace: wvn_slic_workspace_bytes(). */
void wvn_slic_tables(int* g256, int* m9, int* f4096);
int wvn_slic_geometry(int h, int w, int num_components, int* grid_interval, int* nx, int* ny);
size_t wvn_slic_workspace_bytes(int batch, int h, int w, int num_components);
int wvn_slic(const float* img, int batch, int h, int w, int num_components, float compactness, int iters,
             const int* lut_g, const int* lut_m, const int* lut_f, long long* labels, void* workspace, void* stream);

/* Footprint projection + rasterisation — replaces ImageProjector.project_and_render
 * (image_projector/image_projector.py:152-197; its kornia calls transform_points, PinholeCamera.project and
 * draw_convex_polygon are restated, see oracle/image_projector.py) and, when supervision_inout != NULL, the mask update of
 * TraversabilityEstimator.add_supervision_node (traversability_estimator/traversability_estimator.py:281-284):
 *   supervision = fmin(supervision, mask * traversability).
 * K: [batch,4,4] SCALED camera matrices (ImageProjector.__init__ :61-75); pose_camera_in_world: [batch,4,4];
 * points: [batch,n_points,3] convex polygon in the world frame; colors: [batch,3] (color_batched) or [3];
 * traversability: device scalar or NULL (= 1).  Outputs (each may be NULL): masks [batch,3,h,w] fp32 with NaN outside the
 * polygon (and where the colour is 0, as `masks[masks == 0] = nan`); projected [batch,n_points,2] pixel coordinates
 * (NaN for points behind the camera); valid [batch,n_points] uint8 (check_validity :103-124). */
int wvn_project_and_render(const float* K, const float* pose_camera_in_world, const float* points, const float* colors,
                           int color_batched, int batch, int n_points, int h, int w, const float* traversability,
                           float* masks, float* projected, unsigned char* valid, float* supervision_inout, void* stream);

/* ------------------------------------------------------------------------------------------
 * Traversability MLP inference over every pixel — replaces
 *   x = dense_feat[0].permute(1,2,0).reshape(-1, D); prediction = model.forward(Data(x));
 *   out_trav = prediction[...,0]; loss_reco = mse(prediction[:,1:], x).mean(1);
 *   confidence = confidence_generator.inference_without_update(loss_reco)
 * (wild_visual_navigation_ros/scripts/wvn_feature_extractor_node.py:319-370,
 *  model/simple_mlp.py:33-39, utils/confidence_generator.py:182-193).
 * ---------------------------------------------------------------------------------------- */
typedef struct wvn_mlp_infer wvn_mlp_infer_t;
int wvn_mlp_infer_create(int dim, int h1, int h2, int chunk_rows, wvn_mlp_infer_t** out);
void wvn_mlp_infer_destroy(wvn_mlp_infer_t* h);
/* Sizes the per-pixel path's workspaces for a token grid of `tokens_per_frame` patches, so that wvn_mlp_infer_pixels
 * never allocates (call once after create; a later call with a larger grid grows them). */
int wvn_mlp_infer_reserve(wvn_mlp_infer_t* h, int tokens_per_frame);
/* params: flat fp32 buffer in state_dict order layers.0.weight, layers.0.bias, layers.2.weight,
 * layers.2.bias, layers.4.weight, layers.4.bias (device memory). */
int wvn_mlp_infer_set_params(wvn_mlp_infer_t* h, const float* params, void* stream);
/* tokens: [batch, gh*gw, dim] fp32; trav, conf: [batch, out_h, out_w] fp32;
 * cg_mean, cg_std: device scalars of the ConfidenceGenerator. */
int wvn_mlp_infer_pixels(wvn_mlp_infer_t* h, const float* tokens, int batch, int gh, int gw, int out_h, int out_w,
                         const float* cg_mean, const float* cg_std, float std_factor, float* trav, float* conf,
                         void* stream);
/* The same maps straight from the backbone's own bf16 token buffer (the tokens of the last wvn_vit_forward* call, frames
 * [0, batch)): no fp32 -> bf16 re-cast of the tokens.  Needs dim == the backbone's width and the fused geometry. */
int wvn_mlp_infer_pixels_vit(wvn_mlp_infer_t* h, wvn_vit_t* vit, int batch, int out_h, int out_w, const float* cg_mean,
                             const float* cg_std, float std_factor, float* trav, float* conf, void* stream);
/* Same arithmetic on explicit rows x: [rows, dim] fp32 (segment-wise prediction mode). */
int wvn_mlp_infer_rows(wvn_mlp_infer_t* h, const float* x, long long rows, const float* cg_mean,
                       const float* cg_std, float std_factor, float* trav, float* conf, void* stream);

/* ------------------------------------------------------------------------------------------
 * Online train step (fp32) — replaces the body of TraversabilityEstimator.train()
 * (traversability_estimator/traversability_estimator.py:464-477): SimpleMLP.forward,
 * TraversabilityLoss.forward (utils/loss.py:93-160) with the ConfidenceGenerator
 * "latest_measurement" update (utils/confidence_generator.py:78-82), backward, Adam.
 * Three phases so a data-parallel caller can all-reduce between them:
 *   forward_stats -> [all-reduce scalars[0..4] (5 doubles)] -> backward
 *   -> [all-reduce grads (n_params + 1 floats)] -> finalize + adam.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float w_trav, w_reco, std_factor;
  int anomaly_balanced;
  float lr, beta1, beta2, eps;
} wvn_train_config;

size_t wvn_mlp_param_count(int dim, int h1, int h2);
size_t wvn_mlp_train_workspace_bytes(int dim, int h1, int h2, int max_rows);
size_t wvn_mlp_train_scalars_bytes(void);
int wvn_mlp_train_forward_stats(int dim, int h1, int h2, const float* params, const float* x, const float* y,
                                const unsigned char* y_valid, int rows, int max_rows, void* workspace, void* scalars,
                                void* stream);
int wvn_mlp_train_backward(int dim, int h1, int h2, const float* params, const float* x, const float* y,
                           const unsigned char* y_valid, int rows, int max_rows, long long n_total,
                           const wvn_train_config* cfg, void* workspace, void* scalars, float* cg_mean, float* cg_std,
                           float* grads, float* confidence_out, void* stream);
int wvn_mlp_train_apply(int dim, int h1, int h2, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                        long long* step_counter, long long n_total, const wvn_train_config* cfg, void* scalars,
                        void* stream);
/* metrics_out (device, 6 floats): loss_total, loss_trav, loss_reco, loss_trav_confidence, cg_mean, cg_std */
int wvn_mlp_train_read_metrics(const void* scalars, float* metrics_out, void* stream);
/* fp32 forward only: out [rows, 1+dim] (column 0 through the sigmoid), SimpleMLP.forward. */
int wvn_mlp_forward_f32(int dim, int h1, int h2, const float* params, const float* x, int rows, float* h1_buf,
                        float* h2_buf, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused online train step — the same arithmetic as the three-phase entry points above in FOUR kernels, with the
 * row compaction (`feat[seg_mask]`, wvn_feature_extractor_node.py:324-327 / nodes.py:199-241) and, for data-parallel
 * runs, the two all-reduces inside the library (SURVEY.md §8b "wvn_mlp_train_step(..., ncclComm_t or NULL, ...)",
 * §8e).  No host synchronisation, no allocation after create.
 *   x: [groups, rows_per_group, dim] fp32, PADDED per group; n_rows: [groups] int32 (device) live rows per group, or
 *   NULL when every row is live;  y [fp32] / y_valid [uint8] / confidence_out [fp32] are indexed by the COMPACTED row
 *   number (live rows of group 0, then group 1, ...);  metrics_out (device, 6 floats, may be NULL): loss_total,
 *   loss_trav, loss_reco, loss_trav_confidence, cg_mean, cg_std.  params / exp_avg / exp_avg_sq: flat fp32 buffers in
 *   state_dict order (torch.optim.Adam state), step_counter: device int64.
 * phase_mask: 7 = whole step; 1 / 2 / 4 = forward+stats / backward+weight-gradients / metrics+Adam separately (a caller
 * without a library communicator all-reduces `scalars` (6 doubles) after phase 1 and `grads` (n_params + 1) after 2).
 * ---------------------------------------------------------------------------------------- */
typedef struct wvn_mlp_trainer wvn_mlp_trainer_t;
/* scalars: optional caller-owned device buffer of wvn_mlp_trainer_scalars_bytes(); grads: optional caller-owned device
 * buffer of n_params + 1 floats (NULL: the trainer allocates them with its workspace). */
size_t wvn_mlp_trainer_scalars_bytes(void);
int wvn_mlp_trainer_create(int dim, int h1, int h2, int max_rows, const wvn_train_config* cfg, void* scalars,
                           float* grads, wvn_mlp_trainer_t** out);
void wvn_mlp_trainer_destroy(wvn_mlp_trainer_t* t);
/* Library-owned NCCL communicator (libnccl.so.2 is resolved from the running process): rank 0 fills a 128-byte id with
 * wvn_comm_unique_id, the caller broadcasts it by any means, every rank calls wvn_mlp_trainer_init_comm. */
int wvn_comm_unique_id(void* id128);
int wvn_mlp_trainer_init_comm(wvn_mlp_trainer_t* t, const void* id128, int rank, int world);
/* ConfidenceGenerator method of the fused step (utils/confidence_generator.py:49-76): 0 latest_measurement (default),
 * 1 running_mean (:94-115), 2 kalman_filter (:131-145 with utils/kalman_filter.py:78-111, D = 1, F = H = 1; kf_proc_cov /
 * kf_meas_cov = its Q / R), 3 moving_average (:117-129, window of 5 steps kept inside the trainer).  The state the
 * reference keeps in module parameters is read and updated in place through these device pointers — var (1,1) fp32;
 * running_n / running_sum / running_sum_of_squares (1,) fp64 — each may be NULL (the trainer then keeps a private copy). */
int wvn_mlp_trainer_set_confidence(wvn_mlp_trainer_t* t, int method, float* var, double* running_n, double* running_sum,
                                   double* running_sum_of_squares, float kf_proc_cov, float kf_meas_cov);
int wvn_mlp_train_step(wvn_mlp_trainer_t* t, float* params, float* exp_avg, float* exp_avg_sq, long long* step_counter,
                       const float* x, int groups, int rows_per_group, const int* n_rows, const float* y,
                       const unsigned char* y_valid, float* cg_mean, float* cg_std, float* confidence_out,
                       float* metrics_out, int phase_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WVN_B200_H_ */
