/*
 * mdb200.h -- C ABI of libmdb200.so, the B200-native batch detection engine that drops in behind
 * MegaDetector's PyTorch detector.
 *
 * The reference has no FFI: its seam is the Python duck type returned by
 *   run_detector.load_detector()                      megadetector/detection/run_detector.py:601-681
 * i.e. PTDetector                                      megadetector/detection/pytorch_detector.py:739-1482
 * The Python mirror of that class (megadetector_b200/detector.py) binds exactly the entry points declared
 * here through ctypes; INTEGRATION.md shows the reference-side hook.  Each function cites the reference
 * code it replaces.
 *
 * Conventions: every function returns MDB_OK (0) or a negative error; no exception crosses the
 * boundary; the caller owns all host buffers; the library owns device memory inside the opaque handle;
 * one handle per GPU; a handle is not thread-safe, different handles are independent.  There is no CPU
 * fallback: without a CUDA device mdb_create() fails with MDB_ERR_CUDA.
 */
#ifndef MDB200_H_
#define MDB200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MDB_API __attribute__((visibility("default")))
#else
#define MDB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_OK 0
#define MDB_ERR_CUDA (-1)
#define MDB_ERR_ARG (-2)
#define MDB_ERR_STATE (-3)
#define MDB_ERR_NOMEM (-4)

/* per-image status written by the infer calls */
#define MDB_IMG_OK 0
#define MDB_IMG_BAD_INPUT 1   /* maps to FAILURE_IMAGE_OPEN, run_detector.py:55 */
#define MDB_IMG_INFER_FAILED 2 /* maps to FAILURE_INFER,      run_detector.py:56 */

#define MDB_MAX_DET_CAP 300 /* nms(max_det=300), pytorch_detector.py:502 */

typedef struct mdb_engine mdb_engine;

/* YOLOv5-P6 family member.  yolov5x6 / MDv5: ch={80,160,320,640,960,1280},
 * n_backbone={4,8,12,4,4}, n_head=4, nc=3. */
typedef struct mdb_arch {
  int32_t ch[6];
  int32_t n_backbone[5];
  int32_t n_head;
  int32_t nc;
  float anchors_px[4][3][2]; /* Detect.anchors * stride, (w,h) in letterbox pixels */
} mdb_arch;

typedef struct mdb_config {
  int32_t struct_size; /* sizeof(mdb_config), checked */
  int32_t device;      /* CUDA ordinal ("cuda:0" hard-coded at pytorch_detector.py:863-864) */
  int32_t max_batch;   /* images per forward pass */
  int32_t img_size;    /* 1280: default_image_size, pytorch_detector.py:798-810 */
  int32_t stride;      /* 64:   letterbox_stride,   pytorch_detector.py:826-843 */
  int32_t letterbox_auto;    /* 1 in "classic" mode, pytorch_detector.py:1064-1068 */
  int32_t letterbox_scaleup; /* 1 in "classic" mode */
  float iou_thres;     /* 0.45 classic / 0.6 otherwise, pytorch_detector.py:1318-1321 */
  int32_t max_det;     /* 300 */
  /* 0: "classic" (letterbox resize + auto pad), 1: "modern" (pytorch_detector.py:1036-1101: long side resized to img_size with
   * cv2.INTER_AREA when shrinking / INTER_LINEAR when growing, then padded to ceil(side / stride + 0.5) * stride, boxes
   * mapped back with ratio_pad).  letterbox_auto / letterbox_scaleup are ignored in mode 1. */
  int32_t compat_mode;
  int32_t use_ceil_for_resize; /* "use_ceil_for_resize" in compatibility_mode, pytorch_detector.py:1038,1052-1057 */
  mdb_arch arch;
} mdb_config;

/* A named fp32 tensor in the fused-yolov5 state-dict naming and PyTorch layout
 * (conv weight [Cout][Cin][kh][kw], bias [Cout]); see oracle/yolov5.py for the names. */
typedef struct mdb_tensor_desc {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} mdb_tensor_desc;

/* ---- life cycle: replaces PTDetector.__init__/_load_model, pytorch_detector.py:745-959 ---- */
MDB_API int mdb_create(mdb_engine** out, const mdb_config* cfg);
MDB_API int mdb_load_weights(mdb_engine* e, const mdb_tensor_desc* tensors, int32_t n);
MDB_API void mdb_destroy(mdb_engine* e);
/* last error text of this handle (or of the failed mdb_create when e == NULL) */
MDB_API const char* mdb_last_error(const mdb_engine* e);

/* ---- the hot path: replaces PTDetector.generate_detections_one_batch for raw images,
 *      pytorch_detector.py:1124-1252 = preprocess_image (:964) + _process_batch_group (:1257-1424)
 *      up to and including scale_coords(...).round() (:1379).
 * rgb_hwc[i]   : uint8 RGB, h[i] x w[i] x 3, row_stride[i] bytes between rows (pageable or pinned host memory)
 * conf_thres   : detection_threshold (1e-5 on the batch path, run_detector_batch.py:751-754)
 * det_out      : [n][max_det][6] = x1,y1,x2,y2 (original-image pixels, rounded half-to-even), conf, class(0-based)
 *                rows sorted by conf descending (ties: anchor index ascending)
 * det_count    : [n]
 * kept_anchor  : [n][max_det] row index into the model's prediction tensor (for bit-exact index parity); may be NULL
 * status       : [n] MDB_IMG_*
 * Images are grouped by letterboxed shape internally (pytorch_detector.py:1226-1233). */
MDB_API int mdb_infer_batch(mdb_engine* e, const uint8_t* const* rgb_hwc, const int32_t* h, const int32_t* w,
                    const int64_t* row_stride, int32_t n, float conf_thres, float* det_out, int32_t* det_count,
                    int32_t* kept_anchor, int32_t* status);

/* Streaming form of mdb_infer_batch (same arguments; the output arrays must stay valid until the batch is done):
 * begin() plans the chunks and queues the first ones, every next() completes the oldest outstanding chunk and
 * returns the caller-array indices whose results (or failure status) are final, queuing further chunks behind it.
 * *n_done == 0 means the batch is finished.  Lets the host format records (pytorch_detector.py:1384-1411) for
 * chunk i while the GPU runs chunk i+1. */
MDB_API int mdb_batch_begin(mdb_engine* e, const uint8_t* const* rgb_hwc, const int32_t* h, const int32_t* w,
                            const int64_t* row_stride, int32_t n, float conf_thres, float* det_out, int32_t* det_count,
                            int32_t* kept_anchor, int32_t* status);
MDB_API int mdb_batch_next(mdb_engine* e, int32_t* done_idx, int32_t cap, int32_t* n_done);

/* Same, for images already letterboxed by a producer worker (dict inputs, run_detector_batch.py:175-183):
 * lb_hwc[i] is lh[i] x lw[i] x 3 (dense), src_h/src_w the original size used by scale_coords. */
MDB_API int mdb_infer_preprocessed(mdb_engine* e, const uint8_t* const* lb_hwc, const int32_t* lh, const int32_t* lw,
                           const int32_t* src_h, const int32_t* src_w, int32_t n, float conf_thres, float* det_out,
                           int32_t* det_count, int32_t* kept_anchor, int32_t* status);

/* JPEG files decoded ON THE GPU (nvJPEG: hardware JPEG engine or GPU-hybrid backend), straight into the buffers K1 reads:
 * replaces visualization_utils.load_image (megadetector/visualization/visualization_utils.py:103-176,306-323: PIL open ->
 * RGB -> EXIF orientation 3/6/8), which run_detector_batch calls per image on the host (run_detector_batch.py:724).
 * jpeg[i] / nbytes[i]: the file's bytes (host memory).  Unreadable streams and 4-component (CMYK) files get
 * MDB_IMG_BAD_INPUT, as PIL-unreadable files do in the reference.  src_hw (may be NULL): [n][2] decoded height, width
 * (after the orientation has been applied) -- the `scaling_shape` the records are normalised by.
 * Decoded pixels differ from libjpeg-turbo's by a few LSB at chroma edges (different IDCT / up-sampling arithmetic). */
MDB_API int mdb_infer_jpeg_batch(mdb_engine* e, const uint8_t* const* jpeg, const size_t* nbytes, int32_t n, float conf_thres,
                                 float* det_out, int32_t* det_count, int32_t* kept_anchor, int32_t* status, int32_t* src_hw);
/* parity hook: decode one JPEG and return the RGB pixels ([h][w][3] u8), hw2 = {h, w}, backend = 1 hardware / 2 GPU hybrid */
MDB_API int mdb_op_decode_jpeg(mdb_engine* e, const uint8_t* jpeg, size_t nbytes, uint8_t* rgb_out, size_t cap, int32_t* hw2,
                               int32_t* backend);

/* mdb_infer_preprocessed with explicit scale_coords parameters per image: gain_pad[i] = {gain, pad_x, pad_y}, the
 * ratio_pad the "modern" mode passes to scale_coords (pytorch_detector.py:1372-1381: gain = resized_h / original_h,
 * pad = letterbox_pad); NULL = derive them from the shapes ("classic"). */
MDB_API int mdb_infer_preprocessed_rp(mdb_engine* e, const uint8_t* const* lb_hwc, const int32_t* lh, const int32_t* lw,
                              const int32_t* src_h, const int32_t* src_w, const float* gain_pad, int32_t n, float conf_thres,
                              float* det_out, int32_t* det_count, int32_t* kept_anchor, int32_t* status);

/* `image_size` argument of generate_detections_one_batch (pytorch_detector.py:1011-1024): long-side target of
 * the letterbox for subsequent calls (default cfg.img_size). */
MDB_API int mdb_set_image_size(mdb_engine* e, int32_t img_size);

/* Test-time augmentation for the following infer calls: `augment=True` of generate_detections_one_batch
 * (pytorch_detector.py:1313 `self.model(batch_tensor, augment=augment)`), i.e. yolov5 `_forward_augment`: three passes at
 * scales 1 / 0.83 (left-right flipped) / 0.67 (bilinear `scale_img`, padded with 0.447 to a multiple of 64), predictions
 * de-scaled, `_clip_augmented`, concatenated (111 768 rows at 768x1280), then one nms() over all of them. */
MDB_API int mdb_set_augment(mdb_engine* e, int32_t on);

/* Letterbox geometry exactly as the reference computes it (pytorch_detector.py:434-452,1104-1109):
 * out[0..7] = new_unpad_w, new_unpad_h, top, bottom, left, right, out_h, out_w ; ratio_pad[0..2] = r, dw, dh */
MDB_API int mdb_letterbox_geometry(const mdb_config* cfg, int32_t h, int32_t w, int32_t* out8, double* ratio_pad3);

/* ---- parity / profiling hooks (not used by the product path) ---- */
#define MDB_FETCH_STEM_INPUT 1 /* K1 output of the last batch: fp16 [n][H/2][W/2][48]            */
#define MDB_FETCH_PRED 2       /* decoded predictions of the last batch: fp32 [n][anchors][5+nc]  */
#define MDB_FETCH_LAYER 3      /* fp16 NHWC output of conv #index (order of oracle.yolov5.conv_specs) */
MDB_API int mdb_debug_fetch(mdb_engine* e, int32_t what, int32_t index, void* host_dst, size_t bytes, int64_t* shape4);

/* Parity harness: while a trace is set, the forward pass runs eagerly and keeps dense copies of the input (which=0),
 * output (1) and residual (2) tensors of the selected steps, taken right before / after each step runs (so in-place steps
 * and the liveness-packed arena do not matter).  names: NULL or "" = off, "*" = every step, else a comma-separated list of
 * step names (conv names of oracle.yolov5.conv_specs; "model.N.cv12.conv" for the fused C3 entry pair; "upsample.13|17|21";
 * "sppf_pool.11").  This is what lets tests compare every one of the 163 convolutions of pytorch_detector.py:1313 with a
 * PyTorch evaluation of the same layer on the engine's own input.  shape5 = N, H, W, C, is_fp32. */
/* Parity harness only, no GPU needed: the addressing tables of K2's 3x3 halo main loop (how the TMA boxes of one
 * 64-channel chunk are laid out in an A stage and where the tcgen05 descriptor of filter tap (kh, kw) starts), so that a test can
 * model which input pixel every MMA row of every tap reads -- the im2col that cuDNN performs inside the reference's
 * `model(x)` (pytorch_detector.py:1313).  mode 0: stride 1, one box per kw shift; 1: stride 1, single box; 2: stride 2,
 * four parity-view boxes.  out64: [0] boxes, [1] bytes per chunk, [2] stage bytes, then per box (4 each) offset, tensor map,
 * parity view (-1: the tensor itself, else py * 2 + px), dx, dy, width, height from [4]; per tap (9 each) byte offset from [32],
 * group stride >> 4 from [41]. */
MDB_API int mdb_debug_halo_layout(int32_t mode, int32_t* out64);
MDB_API int mdb_debug_set_trace(mdb_engine* e, const char* names);
MDB_API int mdb_debug_trace_names(mdb_engine* e, char* names_out, size_t cap, int32_t* n_steps);
MDB_API int mdb_debug_fetch_trace(mdb_engine* e, const char* name, int32_t which, void* host_dst, size_t bytes, int64_t* shape5);

/* device timing of the last infer call, milliseconds: [0]=h2d+preprocess [1]=network [2]=decode+nms [3]=total */
MDB_API int mdb_last_timing(mdb_engine* e, float* ms4);
/* Device-resident throughput measurement (bench.py `value`): re-runs the batch that the last infer call left
 * staged in HBM `iters` times back to back (stages_mask: 1 = K1 preprocess, 2 = network, 4 = decode+NMS+result
 * D2H; 7 = everything) and returns the elapsed device time between CUDA events on the compute stream. */
MDB_API int mdb_bench_resident(mdb_engine* e, int32_t iters, float conf_thres, int32_t stages_mask, float* ms_total);
/* Per-step device time of the staged batch (each step launched `iters` times back to back): ms_out[i], the
 * conv multiply-accumulates of the step macs_out[i] (0 for non-conv steps) and a description in
 * names_out[i * name_stride].  *n_steps receives the number of steps of the plan. */
MDB_API int mdb_profile_steps(mdb_engine* e, int32_t iters, float* ms_out, double* macs_out, char* names_out,
                              int32_t name_stride, int32_t cap, int32_t* n_steps);
/* A group of n <= max_batch images is uploaded and processed in `parts` chunks so that the H2D copy of one
 * chunk overlaps the kernels of the previous one (default 2; 1 = one chunk). */
MDB_API int mdb_set_pipeline_split(mdb_engine* e, int32_t parts);
/* number of kernels the engine launched since creation (bench.py's gpu_launches) */
MDB_API int64_t mdb_launch_count(const mdb_engine* e);

/* ---- single-kernel entry points on HOST buffers, for the parity tests (tests/ -m gpu) ---- */
/* K1: letterbox + pad(114) + /255 + space-to-depth NHWC fp16.  dst: [out_h/2][out_w/2][48] fp16
 * (three column taps of the stem packed into the channel dimension, see csrc/preprocess.cu). */
MDB_API int mdb_op_letterbox(mdb_engine* e, const uint8_t* rgb_hwc, int32_t h, int32_t w, int64_t row_stride, void* dst_fp16,
                     int32_t* out_hw2);
/* K2: one convolution.  x: fp16 NHWC [n][h][w][cin]; wgt: fp32 [cout][cin][k][k]; bias fp32 [cout];
 * residual (may be NULL): fp16 NHWC [n][ho][wo][cout]; y: fp16 (or fp32 when out_fp32) NHWC [n][ho][wo][cout].
 * iters>1 repeats the launch and reports the mean device time in *ms (may be NULL). */
MDB_API int mdb_op_conv2d(mdb_engine* e, const void* x_fp16, int32_t n, int32_t h, int32_t w, int32_t cin, const float* wgt,
                  const float* bias, int32_t cout, int32_t k, int32_t s, int32_t pad, int32_t act,
                  const void* residual_fp16, int32_t out_fp32, void* y, int32_t iters, float* ms);
/* K3: nn.Upsample(scale_factor=2, mode="nearest") (yolov5x6.yaml idx 13/17/21).  src fp16 NHWC [n][h][w][c] -> dst [n][2h][2w][c] */
MDB_API int mdb_op_upsample2x(mdb_engine* e, const void* src_fp16, int32_t n, int32_t h, int32_t w, int32_t c, void* dst_fp16);
/* K4: the three chained MaxPool2d(5,1,2) of SPPF (yolov5 models/common.py) in one pass.  s1 fp16 NHWC [n][h][w][c] ->
 * cat [n][h][w][4c] = [s1 | p1 | p2 | p3] */
MDB_API int mdb_op_sppf_pool(mdb_engine* e, const void* s1_fp16, int32_t n, int32_t h, int32_t w, int32_t c, void* cat_fp16);
/* K5: Detect inference decode (yolov5 models/yolo.py Detect.forward; oracle/yolov5.py::decode_level) of the four raw head
 * outputs raw4[l]: fp32 NHWC [n][ny4[l]][nx4[l]][round_up(3*(5+nc), 4)] with the anchors of the engine's config ->
 * pred fp32 [n][sum_l 3*ny*nx][5+nc] */
MDB_API int mdb_op_decode(mdb_engine* e, const float* const* raw4, const int32_t* ny4, const int32_t* nx4, int32_t n, float* pred);
/* K5-K8: threshold + per-class NMS + top-k on a prediction tensor fp32 [n][anchors][5+nc] (the reference's
 * nms(), pytorch_detector.py:502-610).  det_out boxes stay in letterbox pixels (no rescale). */
MDB_API int mdb_op_nms(mdb_engine* e, const float* pred, int32_t n, int32_t anchors, int32_t nc, float conf_thres,
               double iou_thres, int32_t max_det, float* det_out, int32_t* det_count, int32_t* kept_anchor);

/* ---- host memory helpers (page-locked buffers make the H2D copy asynchronous) ---- */
MDB_API void* mdb_alloc_pinned(size_t bytes);
MDB_API void mdb_free_pinned(void* p);

MDB_API const char* mdb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MDB200_H_ */
