/* vsr_b200 — C ABI of the B200-native STTN inpainting engine.
 *
 * The reference (YaoFANGUK/video-subtitle-remover) has no FFI: its plug-in boundary is a set of Python
 * callables.  Each entry point below names the reference call it replaces (paths relative to the
 * reference root); the ctypes binding that a maintainer adds is in INTEGRATION.md and
 * vsr_b200/_capi.py.  Plain pointers and sizes only — no torch / numpy types cross this boundary.
 *
 * All functions return 0 on success and a negative code on failure; vsr_last_error() returns a
 * human-readable message for the calling thread.  There is no CPU fallback: every compute entry
 * point fails with VSR_ERR_CUDA when no sm_100 device is present.
 */
#ifndef VSR_B200_H
#define VSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSR_OK 0
#define VSR_ERR_ARG (-1)
#define VSR_ERR_CUDA (-2)
#define VSR_ERR_STATE (-3)
#define VSR_ERR_NOMEM (-4)
#define VSR_ERR_RANGE (-5)   /* attention logits beyond the single-pass softmax's range: set option attn_direct = 0 and repeat the call */

typedef struct vsr_sttn vsr_sttn_t;

/* Network geometry + window schedule (backend/inpaint/sttn/auto_sttn.py:64-95 and
 * backend/inpaint/sttn_auto_inpaint.py:38-41). */
typedef struct vsr_sttn_config {
  int32_t model_w, model_h;      /* 640, 120  (sttn_auto_inpaint.py:38) */
  int32_t n_patch;               /* 4 */
  int32_t patch_w[4], patch_h[4];/* (80,15) (32,6) (10,5) (5,3)  (auto_sttn.py:69) */
  int32_t neighbor_stride;       /* config.sttnNeighborStride = 5 (backend/config.py:89) */
  int32_t ref_length;            /* config.sttnReferenceLength = 10 (backend/config.py:91) */
  int32_t mode;                  /* 0 = sttn-auto (sttn_auto_inpaint.py), 1 = sttn-det (sttn_det_inpaint.py): masked encoder
                                    input, low-res composite with the resized mask, whole strip replaced */
} vsr_sttn_config;

const char* vsr_last_error(void);
const char* vsr_version(void);
/* number of CUDA devices with compute capability 10.x visible to the process */
int vsr_device_count(void);

/* ---- engine life cycle: replaces STTNInpaint.__init__ (sttn_auto_inpaint.py:29-41) ------------- */
void vsr_sttn_default_config(vsr_sttn_config* cfg);
/* STTNDetInpaint geometry: 432x240, patches (108,60) (36,20) (18,10) (9,5), mode 1
 * (sttn_det_inpaint.py:33, sttn/network_sttn.py:70). */
void vsr_sttn_det_config(vsr_sttn_config* cfg);
int vsr_sttn_create(vsr_sttn_t** out, int device, const vsr_sttn_config* cfg);
void vsr_sttn_destroy(vsr_sttn_t* h);
/* One call per tensor of ckpt['netG'] (name as in the state dict, e.g.
 * "transformer.3.attention.query_embedding.weight"); fp32, C-contiguous, torch layout
 * [Cout,Cin,kh,kw] / [Cout].  Replaces model.load_state_dict (sttn_auto_inpaint.py:34). */
int vsr_sttn_set_weight(vsr_sttn_t* h, const char* name, const float* data, const int64_t* shape, int ndim);
/* Packs the weights into the device layout (fp16, tap-major K); fails if any tensor is missing. */
int vsr_sttn_finalize_weights(vsr_sttn_t* h);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* STTNInpaint.inpaint (sttn_auto_inpaint.py:122-164): T strip frames [T,model_h,model_w,3] u8 BGR
 * (host) -> comps [T,model_h,model_w,3] fp32 RGB (host) and visits[T] (1 = the comp is the reference's
 * uint8 single-visit case, >1 = float32 blended). */
int vsr_sttn_inpaint_strip(vsr_sttn_t* h, const uint8_t* frames_bgr, int T, float* comps_out, int32_t* visits_out);
/* STTNDetInpaint.inpaint (sttn_det_inpaint.py:124-174): as above plus the resized mask [model_h,model_w] u8
 * (0..255, one mask for all frames as STTNDetInpaint.__call__ builds it); mode 1 engines only. */
int vsr_sttn_inpaint_strip_masked(vsr_sttn_t* h, const uint8_t* frames_bgr, const uint8_t* mask_small, int T, float* comps_out,
                                  int32_t* visits_out);

/* STTNInpaint.__call__ (sttn_auto_inpaint.py:43-97): T frames of H x W x 3 u8 BGR given as host
 * pointers, mask H x W u8 (>127 = subtitle).  frames_out[i] may equal frames_in[i] (in place, the
 * STTNAutoInpaint chunk loop :242-328) or be distinct buffers (the copy made at :58). */
int vsr_sttn_inpaint_frames(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask,
                            uint8_t* const* frames_out);

/* Split form of the call above for device-resident timing (bench.py `value`):
 *   stage  : host frames -> device strips (H2D)            [also computes the strips from the mask]
 *   compute: everything on the device, returns after the work has been enqueued
 *   fetch  : device strips -> host frames (D2H), synchronises                                      */
int vsr_sttn_stage(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask);
int vsr_sttn_compute(vsr_sttn_t* h);
int vsr_sttn_fetch(vsr_sttn_t* h, uint8_t* const* frames_out);
int vsr_sttn_sync(vsr_sttn_t* h);
/* Asynchronous form of vsr_sttn_inpaint_frames for the chunk loop of STTNAutoInpaint.__call__
 * (sttn_auto_inpaint.py:242-328): submit() copies the strips to pinned memory and enqueues H2D, kernels and
 * D2H, returning a ticket (>= 0) without waiting; collect() waits for that chunk and writes the result into
 * frames_out (which may be the submitted frames).  At most two chunks may be in flight; the frames passed
 * to submit() must stay valid and unmodified until their collect(); the mask must produce exactly one strip
 * and may only change while nothing is in flight. */
int64_t vsr_sttn_submit(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask);
int vsr_sttn_collect(vsr_sttn_t* h, int64_t ticket, uint8_t* const* frames_out);
/* Window-level sharding of ONE chunk over `world` GPUs (single-clip strong scaling).  The windows of the chunk's schedule
 * (sttn_auto_inpaint.py:142-146) are dealt round-robin (window w on rank w % world).  Call order on every rank, same arguments:
 *   shard_begin   stages the strips, encodes the frames this rank's windows decode plus the reference frames it is the home of, and packs
 *                 those references' encoder features into its region of the reference exchange buffer `*ref_buf`
 *   -> all-gather of `*ref_buf`: world regions of `*ref_region_bytes` bytes, region r written by rank r — the features of the
 *      reference frames are the only data a window needs from outside its neighbourhood (get_ref_index, :107-120)
 *   shard_windows runs this rank's windows, each into its own slot of the prediction exchange buffer `*pred_buf`
 *   -> all-gather of `*pred_buf` (world regions of `*pred_region_bytes` bytes)
 *   shard_finish  replays the ordered 0.5 / 0.5 blend (:159-162) on all predictions (bit-identical to the unsharded chunk), composites,
 *                 and writes the frames f with f %% world == rank to frames_out[f] (other entries are not touched).
 * The buffers are device memory owned by the engine; the collectives are the caller's (torch.distributed / NCCL on device tensors that
 * vsr_sttn_copy fills from / drains into a rank's region: device-to-device, nothing passes through the host). */
int vsr_sttn_shard_begin(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask, int rank, int world,
                         void** ref_buf, int64_t* ref_region_bytes, void** pred_buf, int64_t* pred_region_bytes);
/* device -> device copy on the engine's stream, completed on return: moves a rank's region of an exchange buffer into / out of the
 * caller's collective buffers (torch tensors for torch.distributed / NCCL) without exposing the engine's memory to another allocator */
int vsr_sttn_copy(vsr_sttn_t* h, void* dst, const void* src, int64_t bytes);
int vsr_sttn_shard_windows(vsr_sttn_t* h);
int vsr_sttn_shard_finish(vsr_sttn_t* h, uint8_t* const* frames_out);
/* Engine options: "attn_direct" (1: single-pass softmax (no row shift) for the attention heads without split-K — no score matrix, no softmax
 * kernel; a job whose logits leave its range fails with VSR_ERR_RANGE and must be repeated with 0), "use_graph" (CUDA graph per chunk). */
int vsr_sttn_set_option(vsr_sttn_t* h, const char* name, int value);
/* CUDA stream of the engine (cudaStream_t as void*) so callers can bracket it with events. */
void* vsr_sttn_stream(vsr_sttn_t* h);
/* kernels launched by this engine since creation (bench.py `gpu_launches`) */
int64_t vsr_sttn_launch_count(vsr_sttn_t* h);
/* In-situ profile of one chunk (bench.py `roofline`): runs the staged job once eagerly (no CUDA graph) with a pair of CUDA events
 * around every launch group, and returns per class the summed device time (ms) and the number of groups.  Classes: */
enum {
  VSR_PROF_CONV3 = 0,      /* transformer-block 3x3 conv 256->256, fp16 out (feed_forward.conv.0, auto_sttn.py:215) */
  VSR_PROF_CONV3_RES = 1,  /* ... with the fp32 residual stream in the epilogue (output_linear :163, feed_forward.conv.2 :217) */
  VSR_PROF_QKV = 2,        /* fused Q,K,V 1x1 projections (:172-174) */
  VSR_PROF_SCORE = 3,      /* Q.K^T (:143) */
  VSR_PROF_SOFTMAX = 4,    /* softmax rows (:144) */
  VSR_PROF_PV = 5,         /* P.V + un-patch (:145, :201-202) */
  VSR_PROF_ENCODER = 6,    /* :75-84 */
  VSR_PROF_DECODER = 7,    /* :87-95 + tanh / quantise / blend (sttn_auto_inpaint.py:150-162) */
  VSR_PROF_GATHER = 8,     /* feats[neighbor_ids + ref_ids] (sttn_auto_inpaint.py:148) */
  VSR_PROF_PREPOST = 9,    /* crop / resize in, resize back / composite (sttn_auto_inpaint.py:269-271, 312-315) */
  VSR_PROF_CLASSES = 10
};
int vsr_sttn_profile(vsr_sttn_t* h, float* ms_by_class, int64_t* scopes_by_class, int n_classes);
/* Time the dominant kernel family of the last compute: fills ms_out[0..n) with CUDA-event durations
 * of n back-to-back launches of the transformer-block 3x3 conv (tcgen05 implicit GEMM) on T frames. */
int vsr_sttn_time_conv(vsr_sttn_t* h, int T, int n, float* ms_out);

/* Role-level wait accounting of the tcgen05 kernels (all zeros unless the library was built with
 * -DVSR_TC_PROFILE): out64[kernel_id*8 + slot], see csrc/tc_common.cuh.  reset != 0 clears the counters. */
int vsr_debug_tc_profile(uint64_t* out64, int reset);

/* Debug/parity hook: copy an internal activation buffer (NHWC, converted to fp32) to the host.
 * name in {e1,e2s,e3,feats16,feats32,xw16,xw32,att16,comps}; contents are those left by the last compute. */
int vsr_sttn_debug_read(vsr_sttn_t* h, const char* name, float* out, int64_t n);

/* ---- integer mask / index path (host C++, bit-exact) ------------------------------------------ */
/* create_mask (backend/tools/inpaint_tools.py:31-47): boxes = n x (xmin,xmax,ymin,ymax). */
int vsr_create_mask(uint8_t* mask, int H, int W, const int32_t* boxes, int n, int deviation);
/* get_inpaint_area_by_mask (inpaint_tools.py:49-242): mask non-zero = subtitle.  areas = up to
 * max_areas x (ymin,ymax,xmin,xmax); returns the count (>= 0) or a negative error. */
int vsr_inpaint_area_by_mask(int W, int H, int h, const uint8_t* mask, int multiple, int32_t* areas, int max_areas);
/* batch_generator (inpaint_tools.py:7-29) on a length: writes batch sizes, returns their count. */
int vsr_batch_sizes(int n_samples, int max_batch_size, int32_t* sizes, int max_sizes);
/* window schedule (sttn_auto_inpaint.py:142-146 + get_ref_index :107-120): for window w of a T-frame
 * chunk writes neighbour ids then reference ids into ids[] and their counts; returns #windows. */
int vsr_window_schedule(int T, int stride, int ref_length, int32_t* ids, int32_t* n_neighbors, int32_t* n_refs, int max_windows,
                        int max_ids_per_window);

/* ---- device-tensor graph runtime: the DBNet text detector ----------------------------------
 * The reference runs `paddleocr.TextDetection.predict` (backend/tools/subtitle_detect.py:43-58) on the CPU;
 * vsr_b200/dbnet.py compiles the same Paddle PIR program (backend/models/V5/ch_det/inference.json) into calls on
 * these entry points.  Tensors are NHWC fp16 device buffers addressed by raw device pointers (uint64). */
typedef struct vsr_rt vsr_rt_t;
int vsr_rt_create(vsr_rt_t** out, int device);
void vsr_rt_destroy(vsr_rt_t* h);
int vsr_rt_alloc(vsr_rt_t* h, int64_t bytes, uint64_t* dev_ptr);             /* zero-initialised, freed with the runtime or by vsr_rt_free */
int vsr_rt_free(vsr_rt_t* h, uint64_t dev_ptr);                                /* waits for the runtime's stream, then releases one vsr_rt_alloc buffer (ProPainter: work buffers of a batch length that propainter_mode, main.py:229-241, no longer feeds) */
int vsr_rt_upload(vsr_rt_t* h, uint64_t dev_ptr, const void* host, int64_t bytes);
int vsr_rt_download(vsr_rt_t* h, uint64_t dev_ptr, void* host, int64_t bytes);
int vsr_rt_copy(vsr_rt_t* h, uint64_t dst, uint64_t src, int64_t bytes);         /* device -> device on the runtime's stream (ProPainter: chunk results of propainter_inpaint.py:254-304 into the whole-sequence buffers) */
int vsr_rt_sync(vsr_rt_t* h);
int64_t vsr_rt_launch_count(vsr_rt_t* h);
/* Scene-cut scores: what the vendored PySceneDetect ContentDetector computes per frame for SubtitleDetect.get_scene_div_frame_no
 * (backend/tools/subtitle_detect.py:158-170; backend/scenedetect/scene_manager.py:929-933, detectors/content_detector.py:25-35,155-186): down-scale
 * by W // 256 like cv2.resize INTER_LINEAR, 8-bit HSV like cv2.cvtColor, sum |delta| of hue / saturation / value against the previous frame.
 * scene_begin fixes the frame size and forgets the previous frame; scene_frames consumes n consecutive decoded BGR frames (host pointers, H x W x 3)
 * and returns three int64 sums per frame (zeros for the very first frame of a sequence) — integers, so the float score and the threshold /
 * min-scene-length logic on the host reproduce the reference bit for bit. */
int vsr_rt_scene_begin(vsr_rt_t* h, int H, int W);
int vsr_rt_scene_frames(vsr_rt_t* h, const uint8_t* const* frames_bgr, int n, int64_t* sums_out);
/* conv2d / depthwise_conv2d / conv2d_transpose of the PIR program with batch-norm and bias already folded into
 * (w, bias): w fp32 in the framework layout ([Cout,Cin/groups,kh,kw]; transposed: [Cin,Cout,2,2]); cin_pitch =
 * channel pitch of the input tensor.  Dense convs with >= 16 input and >= 8 output channels run on the tcgen05
 * implicit-GEMM kernel (stride 2 through space-to-depth), the rest on small direct kernels. */
int vsr_rt_conv_create(vsr_rt_t* h, const float* w, const float* bias, int cout, int cin, int cin_pitch, int kh, int kw, int stride,
                       int pad_t, int pad_l, int dil, int groups, int transposed, int* layer_id);
/* Per-tensor power-of-two scaling keeps un-normalised activations (the LKPAN neck reaches |x| ~ 1e5) inside fp16: a
 * tensor stores value * s; the layer computes out = acc * alpha + bias * bias_scale with alpha = s_out / s_in and
 * bias_scale = s_out (1, 1 for unscaled tensors).  Every scaled store that leaves fp16 raises the overflow flag. */
/* a dense stride-1 conv whose fp32 weights are kept as hi + lo fp16 halves (every tap issued twice): ~22-bit weights at fp16 activations,
 * for layers whose output is limited by weight rounding (the detector's head).  NOT yet run on a B200 (opt-in). */
int vsr_rt_conv_create_split(vsr_rt_t* h, const float* w, const float* bias, int cout, int cin, int cin_pitch, int kh, int kw, int pad_t, int pad_l, int dil,
                             int* layer_id);
int vsr_rt_conv(vsr_rt_t* h, int layer_id, uint64_t in_ptr, int T, int H, int W, uint64_t out_ptr, int out_pitch, int out_coff, int relu,
                float alpha, float bias_scale);
/* vsr_rt_conv whose output tensor [out_h, out_w] is the window starting at (crop_t, crop_l) of the computed H x W grid:
 * a dense conv over a reflect-padded input keeps only the interior (LAMA's padding_mode='reflect', ReflectionPad2d). */
int vsr_rt_conv_ex(vsr_rt_t* h, int layer_id, uint64_t in_ptr, int T, int H, int W, uint64_t out_ptr, int out_pitch, int out_coff, int relu,
                   float alpha, float bias_scale, int crop_t, int crop_l, int out_h, int out_w);
/* ---- PP-OCRv5 mobile detector (backend/models/V5/ch_det_fast): hardswish + the exported learnable scalar affine, and the
 * squeeze-and-excitation gate (global average pool -> 1x1 conv -> relu -> 1x1 conv -> hardsigmoid), which is applied with
 * vsr_rt_elementwise op 4 (per-channel scale = gate).  residual != 0 gives 1 + gate (RSELayer: x + x * gate). */
int vsr_rt_hswish_affine(vsr_rt_t* h, uint64_t in, uint64_t out, int64_t n_elems, float inv_scale_in, float a, float c);
int vsr_rt_se_create(vsr_rt_t* h, const float* w1, const float* b1, const float* w2, const float* b2, int C, int mid, float slope, float offset,
                     int residual, int* se_id);     /* w1 [mid][C], w2 [C][mid] */
int vsr_rt_se_gate(vsr_rt_t* h, int se_id, uint64_t x, int64_t pixels, int cp, float inv_scale, uint64_t gate_dev);   /* gate: fp32 [cp] */
/* ---- RAFT (backend/inpaint/video/raft/*.py as RAFT_bi calls it, flow_comp_raft.py:39-55): operators beyond the convs.
 * STATUS: checked against the CPU stand-in of the runtime only — not yet run on a B200. */
int vsr_rt_pp_frames(vsr_rt_t* h, const uint8_t* const* frames_bgr, int T, int H, int W, uint64_t out);   /* -> fp16 [T,H,W,8] RGB in [-1,1] */
int vsr_rt_instnorm(vsr_rt_t* h, uint64_t x, int N, int64_t pixels, int cp, int relu, uint64_t out);       /* nn.InstanceNorm2d [+ ReLU] */
int vsr_rt_context_split(vsr_rt_t* h, uint64_t x, int64_t pixels, uint64_t net, int pitch_net, uint64_t inp, int pitch_inp);  /* tanh | relu, raft.py:116-118 */
/* all-pairs correlation / sqrt(C) (corr.py:52-60) as a tensor-core GEMM: out[p1][p2], one row of `out_pitch` halves per source pixel */
int vsr_rt_corr_volume(vsr_rt_t* h, uint64_t fmap1, uint64_t fmap2, int hh, int ww, int C, uint64_t out, int out_pitch);
int vsr_rt_corr_pool(vsr_rt_t* h, uint64_t in, int64_t rows, int h2, int w2, int pitch_in, uint64_t out, int pitch_out);          /* corr.py:24-27 */
/* CorrBlock.__call__ (corr.py:29-50): 4 levels x 9x9 bilinear samples per source pixel at (pixel + flow) / 2^level -> out [pixels][324..] */
int vsr_rt_corr_lookup(vsr_rt_t* h, const uint64_t* level_ptr, const int32_t* level_h, const int32_t* level_w, const int32_t* level_pitch, uint64_t flow32,
                       int hh, int ww, int64_t pixels, uint64_t out, int out_pitch);
int vsr_rt_gru_rh(vsr_rt_t* h, uint64_t r, int pitch_r, uint64_t hsrc, int pitch_h, uint64_t out, int pitch_out, int64_t pixels);   /* sigmoid(r) * h */
int vsr_rt_gru_update(vsr_rt_t* h, uint64_t z, int pitch_z, uint64_t q, int pitch_q, uint64_t hio, int pitch_h, int64_t pixels);   /* update.py:44-56 */
/* flow32 (+)= delta; refresh the fp16 copies (flow16 [P][8]; channels coff, coff+1 of dst_a / dst_b, the GRU input tensors) */
int vsr_rt_flow_update(vsr_rt_t* h, uint64_t flow32, uint64_t delta, int pitch_delta, uint64_t flow16, uint64_t dst_a, uint64_t dst_b, int pitch_ab,
                       int coff, int64_t pixels, int add);
int vsr_rt_convex_upsample(vsr_rt_t* h, uint64_t flow32, uint64_t mask, int pitch_mask, int N, int hh, int ww, uint64_t out32);      /* raft.py:73-84 -> fp32 [N,2,8h,8w] */
/* ---- ProPainter image propagation (propainter.py:107-193 with learnable=False, propainter_inpaint.py:283,311).
 * State tensors are fp16 [T,H,W,8]: channels 0..2 the frame in [-1,1], channel 3 the mask; flows are planar fp32 [2][H][W].
 * STATUS: checked against the CPU stand-in only. */
int vsr_rt_img_prop_step(vsr_rt_t* h, uint64_t prev, uint64_t cur, uint64_t flow_prop, uint64_t flow_check, int H, int W, uint64_t out);
/* prop == 0: state = frame * (1 - m) | m from frames [T,H,W,8] and one device u8 mask [H,W];  else: frame * (1 - m) + prop * m | prop's mask */
int vsr_rt_prop_state(vsr_rt_t* h, uint64_t frames, uint64_t mask_u8, uint64_t prop, int T, int H, int W, uint64_t out);
/* ---- ProPainter flow completion (video/model/recurrent_flow_completion.py): operators beyond the convs.
 * STATUS: checked against the CPU stand-in only. */
int vsr_rt_rfc_input(vsr_rt_t* h, uint64_t flow32, uint64_t mask_u8, int N, int H, int W, int reverse, uint64_t out);      /* [f*(1-m), m] -> fp16 [N,H,W,8] */
int vsr_rt_pad_replicate(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out, int OH, int OW, int top, int left);
int vsr_rt_leaky_relu(vsr_rt_t* h, uint64_t x, int64_t n_elems, float slope);                                               /* in place */
int vsr_rt_temporal_taps(vsr_rt_t* h, uint64_t in, int T, int64_t pixels, int cp_in, uint64_t out, int cp_out);             /* frames t-2, t, t+2 side by side */
/* gather half of torchvision.ops.deform_conv2d (3x3, pad 1, modulated, G offset groups): cols [pixels][9*C]; the 1x1 conv over cols finishes it */
int vsr_rt_deform_cols(vsr_rt_t* h, uint64_t xa, int pitch_a, int Ca, uint64_t xb, int pitch_b, int C, int G, uint64_t om, int pitch_om, float max_residue,
                       uint64_t flow32, int H, int W, int64_t pixels, uint64_t cols, int pitch_cols);
int vsr_rt_rfc_combine(vsr_rt_t* h, uint64_t pred, int pitch_pred, uint64_t flow32, uint64_t mask_u8, int N, int H, int W, int reverse, uint64_t out32);
int vsr_rt_upsample2x_bilinear(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out);   /* F.interpolate(x2, bilinear, align_corners=True) */
/* ---- ProPainter generator, front half (propainter.py:321-378: encoder input, 1/4 flows and masks, the condition tensor of
 * the learnable feature propagation).  STATUS: checked against the CPU stand-in only. */
int vsr_rt_gen_input(vsr_rt_t* h, uint64_t state, uint64_t mask_u8, uint64_t ids_dev, int n, int H, int W, uint64_t out);       /* [rgb, m_in, m_updated] */
int vsr_rt_flow_down4(vsr_rt_t* h, uint64_t flow32, uint64_t ids_dev, int n, int H, int W, uint64_t out32);                       /* -> fp32 [n*h*w][2] */
int vsr_rt_prop_masks(vsr_rt_t* h, uint64_t gen_in, int n, int H, int W, uint64_t out);                                           /* -> fp16 [n*h*w][8] (m_in, m_upd) */
int vsr_rt_featprop_cond(vsr_rt_t* h, uint64_t prop, uint64_t cur, int C, uint64_t flow_prop, uint64_t flow_check, uint64_t masks, int H, int W, uint64_t cond,
                         int pitch);
int vsr_rt_write_extra(vsr_rt_t* h, uint64_t src, uint64_t dst, int pitch, int coff, int nch, int64_t pixels);
/* ---- ProPainter generator, back half (sparse_transformer.py): soft split / composition, layer norm, pooled tokens, the sparse window
 * attention core, the u8 conversion of the decoder output.  STATUS: checked against the CPU stand-in only. */
int vsr_rt_unfold7s3(vsr_rt_t* h, uint64_t in, int n, int hh, int ww, int C, uint64_t out, int pitch, int gelu);       /* nn.Unfold(7,3,3), tap-major */
int vsr_rt_fold7s3(vsr_rt_t* h, uint64_t tok, int n, int hh, int ww, int C, int pitch, int normalise, uint64_t out);   /* F.fold(7,3,3) [/ overlap count] */
int vsr_rt_layernorm(vsr_rt_t* h, uint64_t x, int64_t tokens, int C, uint64_t gamma_dev, uint64_t beta_dev, uint64_t out);
int vsr_rt_pool4(vsr_rt_t* h, uint64_t x, int n, int H, int W, int C, uint64_t w_dev, uint64_t b_dev, uint64_t out);   /* depthwise 4x4 stride 4 */
int vsr_rt_window_attention(vsr_rt_t* h, uint64_t q, uint64_t k, uint64_t v, uint64_t kp, uint64_t vp, int T, int Hn, int Wn, int C, int ph, int pw,
                            uint64_t valid_ind_dev, int n_valid, uint64_t t_ind_dev, int n_tind, uint64_t win_masked_dev, uint64_t out);
int vsr_rt_pred_to_rgb8(vsr_rt_t* h, uint64_t x, int cp, int64_t pixels, uint8_t* host_out);                           /* trunc((tanh(x)+1)/2*255) */
/* ---- LAMA: what the TorchScript big-lama forward needs beyond the detector's operators ---- */
/* out[OH,OW] = in[H,W] shifted by (top, left), reflected at the borders (reflect = 1) or zero filled (0) */
int vsr_rt_pad(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out, int OH, int OW, int top, int left, int reflect);
/* zero insertion [2H, 2W]: a stride-2 transposed conv is a dense conv of it with the flipped kernel */
int vsr_rt_zero_upsample2x(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out);
/* out[p][0..channels) = f(a[p][..]*alpha + b[p][..]*beta) on channel slices (pointers already offset) with their own pitches */
int vsr_rt_add_slices(vsr_rt_t* h, int relu, uint64_t a, int pitch_a, uint64_t b, int pitch_b, uint64_t out, int pitch_out, int channels,
                      int64_t pixels, float alpha, float beta);
/* residual stream with an fp32 master: x32 (+)= y16, x16 = half(x32); init != 0 starts x32 from x16 (FFCResnetBlock's id + x) */
int vsr_rt_residual_add(vsr_rt_t* h, uint64_t x32, uint64_t y16, uint64_t x16, int64_t n_elems, int init);
/* FourierUnit (ffc.py): rfftn(norm='ortho') of C channels of in [H,W,cp_in] -> out [H, W/2+1, 2C] with (re, im) interleaved
 * on the channel axis; the inverse takes that layout back to [H,W,cp_out] (C real channels).  cuFFT, fp32 inside; T images
 * ([T,H,W,..] tensors) go through one batched plan. */
int vsr_rt_fft_r2c(vsr_rt_t* h, uint64_t in, int T, int H, int W, int C, int cp_in, uint64_t out);
int vsr_rt_fft_c2r(vsr_rt_t* h, uint64_t in, int T, int H, int W, int C, uint64_t out, int cp_out);
/* lama_util.py:12-80 + the head of big-lama's forward: host u8 image [h,w,3] and mask [h,w] -> NHWC fp16 [H,W,cp]
 * (img/255*(1-m), m), symmetric padding up to H x W.  The image and mask stay staged on the device in `slot` (one per image
 * of a batch; `out` / `pred` point at that image of the batched tensor) for vsr_rt_lama_output:
 * m*pred + (1-m)*img/255 -> trunc(clip(.*255)) -> host u8 [h,w,3] (lama_inpaint.py:25-27). */
int vsr_rt_lama_input(vsr_rt_t* h, const uint8_t* img, const uint8_t* mask, int ih, int iw, uint64_t out, int H, int W, int cp, int slot);
int vsr_rt_lama_output(vsr_rt_t* h, uint64_t pred, int W, int cp, float inv_scale, int ih, int iw, int slot, uint8_t* out);
/* max |x| of a tensor (inf when it holds a non-finite value): calibration of the scales.  Synchronises. */
int vsr_rt_absmax(vsr_rt_t* h, uint64_t dev_ptr, int64_t n_elems, float* out);
/* host[i] = tensor[i][channel] * mul as fp32, i < pixels (the probability map: one channel of a 64-pitch tensor). */
int vsr_rt_download_channel(vsr_rt_t* h, uint64_t dev_ptr, int64_t pixels, int cp, int channel, float mul, float* host);
/* reads and clears the overflow flag.  Synchronises. */
int vsr_rt_overflow(vsr_rt_t* h, int* raised);
/* Record the launches issued between begin and end into a CUDA graph and replay them with one call; every buffer a
 * recorded launch touches must already exist (run the sequence once before capturing it). */
int vsr_rt_capture_begin(vsr_rt_t* h);
int vsr_rt_capture_end(vsr_rt_t* h, int* graph_id);
int vsr_rt_graph_launch(vsr_rt_t* h, int graph_id);
int vsr_rt_graph_destroy(vsr_rt_t* h, int graph_id);
/* op: 0 a*alpha+b*beta, 1 relu, 2 relu(a*alpha+b*beta), 3 sigmoid(a*alpha), 4 per-channel affine, 5 affine+relu,
 * 6 a*alpha+beta, 7 (a+b)*alpha */
int vsr_rt_elementwise(vsr_rt_t* h, int op, uint64_t a, uint64_t b, uint64_t out, int64_t n_elems, int cp, uint64_t scale_dev,
                       uint64_t shift_dev, float alpha, float beta);   /* scale/shift: device fp32 [cp] (op 4, 5) */
int vsr_rt_upsample_nearest(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, int scale, uint64_t out, int out_pitch, int out_coff);
int vsr_rt_maxpool2x2s1(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out);
int vsr_rt_copy_channels(vsr_rt_t* h, uint64_t src, int src_pitch, uint64_t dst, int dst_pitch, int dst_coff, int channels, int64_t pixels);
/* DetResizeForTest + NormalizeImage (inference.yml:22-40): BGR u8 host image -> NHWC fp16 [dh,dw,cp] on the device */
int vsr_rt_det_preprocess(vsr_rt_t* h, const uint8_t* bgr, int sh, int sw, uint64_t out, int dh, int dw, int cp);

/* ---- operator-level entry points (parity tests call the kernels in isolation) ------------------ */
/* cv2.resize(src_u8 HxWx3, (dw,dh)) INTER_LINEAR on the device (sttn_auto_inpaint.py:72,270). */
int vsr_op_resize_u8(int device, const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw);
/* NHWC conv on the tcgen05 implicit-GEMM kernel: in [T,H,W,Cin] fp32 (host), weight [Cout,Cin,k,k]
 * fp32, k in {1,3}, stride 1, padding = dilation*(k/2); flags bit0 = LeakyReLU(0.2), bit1 = add
 * residual `res` [T,H,W,Cout] fp32.  out [T,H,W,Cout] fp32 (host). */
int vsr_op_conv2d(int device, const float* in, int T, int H, int W, int Cin, const float* weight, const float* bias, int Cout,
                  int ksize, int dilation, int flags, const float* res, float* out);
/* conv3x3 stride 2 pad 1 through the space-to-depth route used for encoder.4 (Cin = 64). */
int vsr_op_conv2d_s2(int device, const float* in, int T, int H, int W, int Cin, const float* weight, const float* bias, int Cout,
                     int flags, float* out);
/* Multi-head patch attention (auto_sttn.py:167-203 without the 1x1/3x3 convs): q,k,v [T,H,W,C] fp32
 * (host), n_patch heads of C/n_patch channels with patches (pw[i], ph[i]); out [T,H,W,C] fp32. */
int vsr_op_patch_attention(int device, const float* q, const float* k, const float* v, int T, int H, int W, int C, int n_patch,
                           const int32_t* pw, const int32_t* ph, float* out);
/* F.interpolate(x, scale_factor=2, 'bilinear', align_corners=True) on NHWC (auto_sttn.py:124-126). */
int vsr_op_upsample2x(int device, const float* in, int T, int H, int W, int C, float* out);

#ifdef __cplusplus
}
#endif
#endif /* VSR_B200_H */
