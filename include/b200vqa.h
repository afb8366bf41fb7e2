/*
 * b200vqa.h - C ABI of libb200vqa.so: the ReLaX-VQA feature-extraction hot path on B200 (sm_100a).
 *
 * The reference (xinyiW915/ReLaX-VQA) has no FFI; its boundary is a set of Python functions
 * (SURVEY.md 8(b)).  Each entry point below replaces the arithmetic of one of them and cites it
 * (paths relative to the reference root).  The Python modules in relax_vqa_b200/ keep the
 * reference's names and signatures and call these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host);
 *   - images are HWC uint8, densely packed, batch-major: [B][H][W][3];
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising; return value 0 = ok, negative = B200VQA_E* code;
 *   - the library never falls back to the CPU: without a CUDA device b200vqa_create fails.
 */
#ifndef B200VQA_H
#define B200VQA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VQA_OK 0
#define B200VQA_EINVAL (-1)     /* bad argument (null pointer, non-positive size, unsupported shape) */
#define B200VQA_ECUDA (-2)      /* a CUDA runtime / driver call failed; see b200vqa_last_error */
#define B200VQA_ENOTLOADED (-3) /* weights for the requested network / head were not loaded */
#define B200VQA_ENOMEM (-4)     /* workspace allocation failed */

#define B200VQA_PATCH 16        /* src/main_fragment_layerstack.py:298 */
#define B200VQA_TARGET 224      /* :297 */
#define B200VQA_TOPN 196        /* :299 */
#define B200VQA_RESNET_STACK 13120
#define B200VQA_RESNET_POOL 2051
#define B200VQA_VIT_POOL 2304
#define B200VQA_FEATURES 35203
/* floats per image of b200vqa_resnet50_maps: the 15 hooked activations (64x112x112, 3 x 256x56x56, 4 x 512x28x28,
 * 4 x 1024x14x14, 3 x 2048x7x7), main_fragment_layerstack.py:91-95 */
#define B200VQA_RESNET_MAP_FLOATS 5920768
#define B200VQA_VIT_TOKENS 196
#define B200VQA_VIT_DIM 768

#define B200VQA_FILTER_BILINEAR 0 /* PIL BILINEAR (antialiased), visualise_resnet.py:41 */
#define B200VQA_FILTER_LANCZOS 1  /* PIL LANCZOS, visualise_vit_layer.py:469 */

typedef struct b200vqa_ctx b200vqa_t;

int b200vqa_version(void);
const char* b200vqa_error_string(int code);
/* text of the last CUDA error seen by this thread's most recent failing call */
const char* b200vqa_last_error(void);

/* context: owns weights, TMA descriptors and grow-only workspaces on `device` */
int b200vqa_create(int device, b200vqa_t** out);
int b200vqa_destroy(b200vqa_t* h);

/* ---- frame sampling from raw yuv420p (upstream of A1): the colour conversion ffmpeg / libswscale applies
 * to the frames selected by extract_frames_yuv / extract_frames_residual_yuv (video_frames_extract.py:29-49, :76-100) before
 * they are written as PNG: unscaled yuv420p -> bgr24, BT.601 limited range, chroma not interpolated, 16-bit fixed point
 * (bit-identical to swscale's x86 path).  yuv: [B] frames of H*W luma bytes + two (H/2)*(W/2) chroma planes, densely packed;
 * bgr: [B][H][W][3].  W % 4 == 0, H even. */
int b200vqa_yuv420p_to_bgr(const uint8_t* yuv, int B, int H, int W, uint8_t* bgr, void* stream);

/* ---- A1+A2 (+gray of A5): cv2.absdiff (main_fragment_layerstack.py:302), get_patch_diff
 * (:177-189), cv2.cvtColor BGR2GRAY (:313-314).  frame/next: [B][H][W][3] BGR.
 * sums: [B][H/16][W/16] exact uint32 patch sums of |next-frame| over 3 channels.
 * residual, gray_frame, gray_next may be NULL (skipped). */
int b200vqa_absdiff_patchsum_u8(const uint8_t* frame, const uint8_t* next, int B, int H, int W,
                                uint8_t* residual, uint32_t* sums, uint8_t* gray_frame,
                                uint8_t* gray_next, void* stream);

/* get_patch_diff on an arbitrary 3-channel image (used for the optical-flow image, :319) */
int b200vqa_patchsum_u8(const uint8_t* img, int B, int H, int W, uint32_t* sums, void* stream);

/* ---- A3: selection half of extract_important_patches (:193-195).  Deterministic tie rule:
 * value descending, flat index ascending (np.argsort(-d, kind="stable")), output in raster
 * order.  pos: [B][top_n][2] int32 (y, x), rows >= count[b] are -1.  count: [B]. */
int b200vqa_topk_patches(const uint32_t* sums, int B, int gh, int gw, int top_n, int32_t* pos,
                         int32_t* count, void* stream);

/* ---- A3 copy half + A4: extract_important_patches (:197-210), get_original_frame_patches
 * (:212-230).  ori_frag = patches of `frame`; diff_frag = patches of |next-frame|.
 * Both [B][224][224][3]; either may be NULL.  Cells past count[b] are zero. */
int b200vqa_gather_fragments(const uint8_t* frame, const uint8_t* next, int B, int H, int W,
                             const int32_t* pos, const int32_t* count, int top_n,
                             uint8_t* ori_frag, uint8_t* diff_frag, void* stream);

/* ---- A8: merge_fragments (:242-245) = cv2.addWeighted(a,.5,b,.5,0), round-half-even */
int b200vqa_merge_fragments(const uint8_t* a, const uint8_t* b, size_t nbytes, uint8_t* out,
                            void* stream);

/* ---- A9: PIL Image.resize((224,224), filter) on uint8, bit-exact (visualise_resnet.py:40-47,
 * visualise_vit_layer.py:466-470).  src [B][H][W][3] -> dst [B][224][224][3].  swap_rb != 0
 * reverses the channel order on output (cv2 BGR frame -> the RGB image PIL would open). */
int b200vqa_resize_pil(b200vqa_t* h, const uint8_t* src, int B, int H, int W, int filter,
                       int swap_rb, uint8_t* dst, void* stream);
/* Both filters of the same frames in one call (every full frame goes to the ResNet through BILINEAR and to the ViT
 * through LANCZOS: main_fragment_layerstack.py:283-288): the horizontal passes share one read of the source.
 * Results are those of two b200vqa_resize_pil calls, bit for bit. */
int b200vqa_resize_pil_pair(b200vqa_t* h, const uint8_t* src, int B, int H, int W, int swap_rb,
                            uint8_t* dst_bilinear, uint8_t* dst_lanczos, void* stream);

/* ---- A5: cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0) (:313-315).
 * gray0/gray1: [B][H][W] uint8; flow: [B][H][W][2] float32 (dx, dy). */
int b200vqa_farneback(b200vqa_t* h, const uint8_t* gray0, const uint8_t* gray1, int B, int H,
                      int W, float* flow, void* stream);
/* A5 + the statistics A6 / A7 need, in one call (:313-319): flow as above, minmax [B][2] = min / max of its magnitude
 * (reduced by the launch that writes the flow) and sums [B][H/16][W/16] = the 16x16 patch sums of flow_to_rgb(flow).
 * Results are those of b200vqa_farneback followed by b200vqa_flow_to_rgb(rgb = NULL), bit for bit. */
int b200vqa_farneback_flow_sums(b200vqa_t* h, const uint8_t* gray0, const uint8_t* gray1, int B,
                                int H, int W, float* flow, uint32_t* sums, float* minmax, void* stream);

/* ---- A6: flow_to_rgb (:162-175).  rgb (really BGR, like the reference): [B][H][W][3] or NULL.
 * sums (patch sums of the colour image, A7) may be NULL.  minmax: [B][2] float scratch/out. */
int b200vqa_flow_to_rgb(const float* flow, int B, int H, int W, uint8_t* rgb, uint32_t* sums,
                        float* minmax, void* stream);

/* ---- A7 copy half + A8 fused: gather the flow-colour patches at `pos` (recomputed from the
 * flow, never materialised) and merge with diff_frag.  flow_frag may be NULL. */
int b200vqa_flow_fragment_merge(const float* flow, const float* minmax, int B, int H, int W,
                                const int32_t* pos, const int32_t* count, int top_n,
                                const uint8_t* diff_frag, uint8_t* flow_frag,
                                uint8_t* merged_frag, void* stream);

/* ---- weights.  h_* are HOST pointers to float32 arrays in PyTorch state-dict layout.
 * names/ptrs/numels describe `n` tensors (torchvision resnet50 keys, visualise_resnet.py:21;
 * DINO VisionTransformer keys, visualise_vit_layer.py:152-260).  BN is kept in fp32 as a
 * per-channel scale/shift applied in the GEMM epilogue; matrices are converted to fp16. */
int b200vqa_load_resnet50(b200vqa_t* h, int n, const char* const* names,
                          const float* const* h_ptrs, const int64_t* numels);
int b200vqa_load_vitb16(b200vqa_t* h, int n, const char* const* names,
                        const float* const* h_ptrs, const int64_t* numels);
/* Mlp head (model_regression.py:37-58) + fitted SimpleImputer / MinMaxScaler attributes
 * (demo_test.py:177-180).  in_features must be 35203. */
int b200vqa_load_head(b200vqa_t* h, int in_features, const float* h_fc1_w, const float* h_fc1_b,
                      const float* h_bn_w, const float* h_bn_b, const float* h_bn_mean,
                      const float* h_bn_var, const float* h_fc2_w, const float* h_fc2_b,
                      const float* h_fc3_w, const float* h_fc3_b, const double* h_imputer_mean,
                      const double* h_scaler_scale, const double* h_scaler_min);

/* ---- A10+A12+A13: ResNet-50 forward on 224x224 images with the 15-hook layer-stack
 * average pooling fused into the producing GEMM epilogues (visualise_resnet.py:62-109,
 * main_fragment_layerstack.py:131-149).  img: [B][224][224][3] uint8, channel order given by
 * is_bgr.  stack: [B][13120] or NULL.  pool: [B][2051] or NULL. */
int b200vqa_resnet50_features(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* stack,
                              float* pool, void* stream);

/* ---- A11+A14+A15: ViT-B/16 forward + [mean,max,std] token pooling
 * (visualise_vit_layer.py:447-500, main_fragment_pool.py:124-132).  out: [B][2304]. */
int b200vqa_vitb16_features(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* out,
                            void* stream);

/* ---- boundary fidelity (debug / drop-in path, not the fused fast path): the RAW arrays the reference's per-image
 * extractors return before the host pools them.
 * b200vqa_resnet50_maps: the 15 hooked activations of visualise_resnet.process_video_frame (visualise_resnet.py:62-109)
 * as fp32 (C,H,W) arrays, concatenated in hook order: maps [B][B200VQA_RESNET_MAP_FLOATS].  conv1 is the raw convolution
 * output (hooked before bn1/relu); the avgpool hook of visualise_resnet_layer.py:62-102 is the spatial mean of the last map.
 * b200vqa_vitb16_tokens: norm(x)[:, 1:] of visualise_vit_layer.py:234-239 / :492-500: tokens [B][196][768] fp32. */
int b200vqa_resnet50_maps(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* maps, void* stream);
int b200vqa_vitb16_tokens(b200vqa_t* h, const uint8_t* img, int B, int is_bgr, float* tokens, void* stream);

/* ---- A16: per-video temporal mean of each block and concatenation into [V][35203]
 * (demo_test.py:171-175).  full_* have rows [full_off[v], full_off[v+1]); frag_* rows
 * [pair_off[v], pair_off[v+1]).  Offsets are device int32 arrays of length V+1. */
int b200vqa_temporal_mean_concat(const float* full_stack, const float* full_vit,
                                 const float* frag_stack, const float* frag_pool,
                                 const float* frag_vit_ori, const float* frag_vit_mer,
                                 const int32_t* full_off, const int32_t* pair_off, int V,
                                 float* features, void* stream);

/* ---- A17+A18: imputer + scaler + Mlp.forward (eval).  features [V][35203] -> score [V] */
int b200vqa_head_forward(b200vqa_t* h, const float* features, int V, float* score, void* stream);

/* ---- training of the Mlp head on the device (fp32), the arithmetic of the reference's loop
 * (model_regression.py:292-306 train_one_epoch, :61-89 MAEAndRankLoss, :375-405 SGD momentum + weight decay / AveragedModel,
 * :454-459 update_bn; fine_tune.py:130-193).  The schedule (epochs, cosine LR, SWA start, k-fold, early stopping) stays on the
 * host: relax_vqa_b200/model_regression.py.  h_* = host float32 arrays in state-dict layout; X / y / masks / loss = device.
 * `which`: 0 = the model being trained, 1 = its SWA average (AveragedModel: parameters averaged, buffers copied at creation). */
typedef struct b200vqa_trainer b200vqa_trainer_t;
int b200vqa_trainer_create(b200vqa_t* h, int in_features, int hidden, b200vqa_trainer_t** out);
int b200vqa_trainer_destroy(b200vqa_trainer_t* t);
int b200vqa_trainer_set_params(b200vqa_trainer_t* t, const float* h_fc1_w, const float* h_fc1_b, const float* h_bn_w, const float* h_bn_b,
                               const float* h_bn_mean, const float* h_bn_var, const float* h_fc2_w, const float* h_fc2_b,
                               const float* h_fc3_w, const float* h_fc3_b);
int b200vqa_trainer_get_params(b200vqa_trainer_t* t, int which, float* h_fc1_w, float* h_fc1_b, float* h_bn_w, float* h_bn_b,
                               float* h_bn_mean, float* h_bn_var, float* h_fc2_w, float* h_fc2_b, float* h_fc3_w, float* h_fc3_b);
/* one optimisation step on a batch X [B][in], y [B] (B >= 2): train-mode forward (batch statistics, running statistics updated
 * with momentum 0.1), MAE + rank loss, backward, SGD.  drop1 [B][hidden] / drop2 [B][hidden/2]: keep masks (1 = keep) of the two
 * Dropout layers, ignored when drop_rate == 0.  loss_out: device float (may be NULL). */
int b200vqa_trainer_step(b200vqa_trainer_t* t, const float* X, const float* y, int B, const uint8_t* drop1, const uint8_t* drop2,
                         float drop_rate, float lr, float momentum, float weight_decay, float l1_w, float rank_w, float* loss_out,
                         void* stream);
int b200vqa_trainer_swa_update(b200vqa_trainer_t* t, void* stream);
/* eval-mode forward (running statistics, no dropout): pred [B] device */
int b200vqa_trainer_predict(b200vqa_trainer_t* t, int which, const float* X, int B, float* pred, void* stream);
/* torch.optim.swa_utils.update_bn, one batch: batch_index 0 restarts the cumulative average of the batch statistics */
int b200vqa_trainer_update_bn(b200vqa_trainer_t* t, int which, const float* X, int B, int batch_index, void* stream);

/* ---- building blocks exposed for tests / profiling ------------------------------------ */
/* D[M][N] = A[M][K] * B[N][K]^T (+bias[N]) on tcgen05; A, B fp16 row-major; D fp32. impl: 0
 * tcgen05 1-CTA kernel, 1 SIMT check kernel, 2 tcgen05 2-CTA (cta_group::2) kernel (N % 256 == 0, K % 64 == 0). */
int b200vqa_gemm_f16(b200vqa_t* h, const void* A, const void* B, const float* bias, float* D,
                     int M, int N, int K, int impl, void* stream);
/* number of kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t b200vqa_launch_count(b200vqa_t* h);
/* debug switch: 0 = tcgen05 (default; linear layers on the 2-CTA cta_group::2 kernel), 1 = SIMT check kernels for
 * every GEMM/conv, 2 = tcgen05 with the 1-CTA kernel everywhere (A/B measurements) */
int b200vqa_set_gemm_impl(b200vqa_t* h, int impl);
/* debug switch for the ViT attention (A/B measurements): 0 = tcgen05 / TMEM kernel (default: S = QK^T and O = PV on the 5th-gen
 * tensor cores, P kept in tensor memory; two independent CTAs per SM), 1 = the warp-level mma.sync kernel of round 1,
 * 2 = the tcgen05 kernel as one CTA per SM with two TMEM regions and the S product issued one tile ahead (measured slower) */
int b200vqa_set_attn_impl(b200vqa_t* h, int impl);
/* scheduling knob: persistent tcgen05 GEMM / conv grids of this context occupy at most `sms` SMs (even; 0 = all), leaving
 * the rest of the GPU to kernels running concurrently on other streams (the bandwidth stages of the next batch).  Results
 * do not depend on it (tiles are computed identically whichever CTA takes them). */
int b200vqa_set_gemm_sms(b200vqa_t* h, int sms);
/* profiling: when on, every tcgen05 GEMM/conv launch is bracketed by CUDA events on its stream.
 * b200vqa_profile_read synchronises, returns the summed device time (ms), the launch count and the
 * algorithmic FLOPs (2*M*N*K of the un-padded problems) since the last read, and resets them. */
int b200vqa_set_profiling(b200vqa_t* h, int on);
/* debug switch for the Farneback kernels (A/B measurements): 0 = streaming column-strip iteration and expansion kernels
 * (default; vertical box sums accumulated in f64 like OpenCV), 1 = the same with Kahan-compensated fp32 sums (6 % faster, but
 * a 1e7:1 edge leaves rounding residue in the flat region below it), 2 = the earlier tile kernels
 * (48 x 32 iteration tiles, 64 x 16 expansion tiles) */
int b200vqa_set_flow_impl(b200vqa_t* h, int impl);
int b200vqa_profile_read(b200vqa_t* h, double* gemm_ms, int64_t* gemm_launches, double* gemm_flops);
/* same for the Farneback iteration kernel (k4_flow_iter): device ms, launches, algorithmic bytes (56 B per pixel) */
int b200vqa_profile_read_flow(b200vqa_t* h, double* ms, int64_t* launches, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* B200VQA_H */
