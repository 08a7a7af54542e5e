// Host-side shared declarations for libmdb200 (engine internals; the public ABI is include/mdb200.h).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

namespace mdb {

constexpr int kMaxDetCap = 300;  // nms(max_det=300), pytorch_detector.py:502

struct CudaError {
  std::string msg;
};

#define MDB_CUDA(expr)                                                                                    \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      char _b[512];                                                                                       \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      throw mdb::CudaError{_b};                                                                           \
    }                                                                                                     \
  } while (0)

#define MDB_REQUIRE(cond, text)                                                          \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      char _b[512];                                                                      \
      snprintf(_b, sizeof(_b), "%s (%s) (%s:%d)", text, #cond, __FILE__, __LINE__);      \
      throw mdb::CudaError{_b};                                                          \
    }                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// A channel slice of an NHWC fp16 (or fp32) buffer.
// ---------------------------------------------------------------------------------------------
struct View {
  void* base = nullptr;  // start of the underlying buffer
  int N = 0, H = 0, W = 0;
  int Ctot = 0;   // channels of the underlying buffer (pixel stride in elements)
  int c_off = 0;  // first channel of this view
  int C = 0;      // channels of this view
  bool fp32 = false;
  size_t elem() const { return fp32 ? 4 : 2; }
  void* ptr() const { return static_cast<char*>(base) + static_cast<size_t>(c_off) * elem(); }
};

// ---------------------------------------------------------------------------------------------
// TMA descriptors (tmap.cu)
// ---------------------------------------------------------------------------------------------
// 4-D {C, W, H, N} fp16 activation map over a View; par_y/par_x < 0: dense, else the (par_y, par_x)
// parity sub-lattice of a stride-2 convolution (dims W/2, H/2).  box = {kc, tw, th, 1}.
void make_act_tmap(CUtensorMap* m, const View& v, int kc, int tw, int th, int par_y, int par_x);
// flattened variant for 1x1 convolutions: dims {C, N*H*W, 1, 1}, box {kc, 128, 1, 1}
void make_act_tmap_flat(CUtensorMap* m, const View& v, int kc);
// 2-D {K, Cout_pad} fp16 weight map, box {kc, bn}
void make_wgt_tmap(CUtensorMap* m, const void* w, int K, int cout_pad, int kc, int bn);

// ---------------------------------------------------------------------------------------------
// K2 convolution (conv_host.cu)
// ---------------------------------------------------------------------------------------------
struct ConvOp;  // opaque here
struct ConvDesc {
  View in, out, res;  // res.base == nullptr: no residual
  const void* wgt = nullptr;  // device fp16 [cout_pad][k][k][cin]
  const float* bias = nullptr;  // device fp32 [cout_pad]
  int cout = 0, cout_pad = 0;
  int cin_pad = 0;  // channels per filter tap in the packed weights (>= in.C, zero padded)
  int k = 1, s = 1, pad = 0, act = 1;
  int kw = 0;      // filter width when it differs from k (0: square)
  int pad_x = -1;  // column padding when it differs from pad (-1: same)
};
// opts the conv kernel of the CURRENT device into its full dynamic shared memory (call once per device, after cudaSetDevice)
// addressing tables of the 3x3 halo main loop (conv_host.cu::halo_layout; mirrored into PairParams by conv_plan)
struct HaloLayout {
  int n_boxes;
  uint32_t a_tx, a_stage_bytes;
  uint32_t box_off[4];
  int box_map[4], box_par[4], box_dx[4], box_dy[4], box_w[4], box_h[4];
  uint32_t tap_off[9], tap_sbo[9];
};
HaloLayout halo_layout(int mode);
void conv_init_device();
ConvOp* conv_build(const ConvDesc& d, int num_sms);
void conv_launch(ConvOp* op, int n_img, cudaStream_t st);
void conv_free(ConvOp* op);
double conv_macs_per_image(const ConvOp* op);
int conv_pick_bn(int cout_pad);
// channels per filter tap in the packed weights: C_in rounded up to a multiple of 64 (zero weights; the TMA unit
// zero-fills the activations)
int conv_cin_pad(int cin);

// weight re-layout helpers (host): PyTorch [co][ci][k][k] fp32 -> [co_pad][k][k][ci_pad] fp16
void pack_weights_khwc(const float* w, int co, int ci, int k, int co_pad, int ci_pad, std::vector<__half>& out);
// stem: [co][3][6][6] -> 3x1 conv over K1's tap-packed space-to-depth image (64 channels = 3 kw taps x 16)
void pack_weights_stem_s2d(const float* w, int co, int co_pad, int ci_pad, std::vector<__half>& out);

// ---------------------------------------------------------------------------------------------
// K1 letterbox / preprocess (preprocess.cu)
// ---------------------------------------------------------------------------------------------
struct LetterboxGeom {
  int src_h, src_w;
  int new_w, new_h;  // new_unpad ("classic") / size after the pre-resize ("modern")
  int top, bottom, left, right;
  int out_h, out_w;
  double ratio, dw, dh;
  int interp = 0;    // 0: cv2.INTER_LINEAR (or no resize), 1: cv2.INTER_AREA ("modern" down-scaling)
  // scale_coords inputs (yolov5 utils.general): classic -> gain/pad derived from the shapes; modern -> ratio_pad given
  double sc_gain = 0, sc_pad_x = 0, sc_pad_y = 0;
};
LetterboxGeom letterbox_geometry(int h, int w, int img_size, int stride, bool autopad, bool scaleup);
// "modern" compatibility mode (pytorch_detector.py:1036-1101): long side resized to img_size (INTER_AREA when shrinking,
// INTER_LINEAR when growing, int() or ceil() of the scaled sides), then padded to ceil(side / stride + 0.5) * stride
LetterboxGeom letterbox_geometry_modern(int h, int w, int img_size, int stride, bool use_ceil);
// coefficient tables of cv::resize(INTER_AREA) for u8 (cv::computeResizeAreaTab): CSR over destination indices
void area_tables(int ssize, int dsize, std::vector<int32_t>& first, std::vector<int32_t>& si, std::vector<float>& alpha);
// coefficient tables of cv::resize(INTER_LINEAR) for u8: ofs[d], coef[2*d], coef[2*d+1]
void linear_tables(int ssize, int dsize, bool vertical, std::vector<int32_t>& ofs, std::vector<int16_t>& coef);

struct PreImage {  // one per image, lives in device memory
  const uint8_t* src;  // device, rows dense or strided
  long long src_stride;
  int src_h, src_w;
  int new_w, new_h, top, left;
  const int32_t* xofs;
  const int16_t* xcoef;
  const int32_t* yofs;
  const int16_t* ycoef;
  // INTER_AREA ("modern" mode, mode != 0): taps of destination column / row d are [first[d], first[d+1])
  int mode = 0;        // 0 linear, 1 area (generic, float taps), 2 area with integer scales (cv::ResizeAreaFast)
  int iscale_x = 0, iscale_y = 0;
  const int32_t* xa_first = nullptr;
  const int32_t* xa_si = nullptr;
  const float* xa_alpha = nullptr;
  const int32_t* ya_first = nullptr;
  const int32_t* ya_si = nullptr;
  const float* ya_alpha = nullptr;
};
void preprocess_launch(const PreImage* d_imgs, int n, int out_h, int out_w, __half* dst, float max_scale,
                       cudaStream_t st);
void preprocess_init_constants();
void preprocess_init_device();  // per device: constants + the kernel's dynamic shared-memory opt-in

// ---------------------------------------------------------------------------------------------
// JPEG decode on the GPU (jpeg.cu): nvJPEG, loaded with dlopen at first use
// ---------------------------------------------------------------------------------------------
struct JpegDecoder;
struct JpegImage {  // decoded, orientation-corrected interleaved RGB image in device memory (dense rows); ptr == nullptr: unreadable
  const uint8_t* ptr;
  int h, w;
};
JpegDecoder* jpeg_create();
void jpeg_destroy(JpegDecoder* d);
int jpeg_decode_batch(JpegDecoder* d, const uint8_t* const* jpeg, const size_t* nbytes, int n, JpegImage* out, cudaStream_t st);
int jpeg_exif_orientation(const uint8_t* p, size_t n);

// ---------------------------------------------------------------------------------------------
// K3/K4/K5 (aux.cu)
// ---------------------------------------------------------------------------------------------
void upsample2x_launch(const View& src, const View& dst, int n_img, cudaStream_t st);
void sppf_pool_launch(const View& s1, const View& p1, const View& p2, const View& p3, int n_img, cudaStream_t st);
struct DecodeLevel {
  const float* raw;  // fp32 [n][ny][nx][c_stride]
  int ny, nx, c_stride;
  float stride;
  float anchors[3][2];
  int row_off;
};
struct DecodePass {  // one pass of test-time augmentation: where its rows go in the concatenated tensor and how to de-scale them
  int out_base, out_rows;
  float scale, flip_w;
};
// thr / cand_*: when cand_keys != nullptr the kernel also runs the candidate pass (K6, nms.cu) on the rows it writes
void decode_launch(const DecodeLevel* lv4, int n_img, int no, int rows_per_img, float* pred, cudaStream_t st,
                   float thr = 0.0f, unsigned long long* cand_keys = nullptr, int32_t* cand_counts = nullptr, int cand_cap = 0,
                   const DecodePass* pass = nullptr);
// TTA input scaling (yolov5 scale_img) from one plan's stem-input tensor into another's
void tta_scale_launch(const __half* src, int H, int W, __half* dst, int Hd, int Wd, double ratio, bool flip, int n_img,
                      cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// K6-K8 (nms.cu)
// ---------------------------------------------------------------------------------------------
struct NmsImageParams {  // scale_coords inputs, one per image (device array); gain == 0: no rescale
  float gain, pad_x, pad_y, max_x, max_y;
};
struct NmsScratch {
  // per image: `cap` 64-bit candidate keys (~bits(conf) << 32 | anchor), double-buffered for the radix passes of long
  // lists, and the candidate count
  unsigned long long* key[2];
  int32_t* count;
  long long* dbg;  // 8 counters per image (phase cycles, -DMDB_NMS_TIMERS builds only)
  int cap;
};
size_t nms_scratch_bytes(int n_img, int anchors);
void nms_bind_scratch(NmsScratch* s, void* mem, int n_img, int anchors);
void nms_init_device();  // per device: dynamic shared-memory opt-in of the per-image kernel
// K6 over a given prediction tensor (the engine runs K6 inside the decode kernel instead): fills scratch.key[0] / count
void nms_candidates_launch(const float* pred, int n_img, int anchors, int no, float conf_thres, const NmsScratch& scratch,
                           cudaStream_t st);
// K7 + K8 on the candidate lists in `scratch`
void nms_launch(const float* pred, int n_img, int anchors, int no, float conf_thres, float iou_thres_floor,
                int max_det, const NmsImageParams* d_params, const NmsScratch& scratch, float* det_out,
                int32_t* det_count, int32_t* kept_anchor, cudaStream_t st);
// largest float <= t  (the CPU torchvision kernel compares float IoU against a double threshold)
float float_floor_of_double(double t);

}  // namespace mdb
