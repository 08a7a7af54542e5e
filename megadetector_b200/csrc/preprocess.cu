// K1: fused letterbox-resize + pad(114) + /255 + HWC->NHWC(space-to-depth) fp16.
//
// Replaces, for one image:
//   * the third-party yolov5 `letterbox()` (cv2.resize INTER_LINEAR + cv2.copyMakeBorder 114) called at
//     megadetector/detection/pytorch_detector.py:1104-1109
//   * HWC->CHW, ascontiguousarray, stack, .float(), /= 255.0 at pytorch_detector.py:1283-1306
// The arithmetic is OpenCV's 8-bit fixed-point bilinear (11-bit coefficients), reproduced exactly:
// coefficient tables are built on the host with the same float/double operations cv::resize uses, the
// kernel is pure integer.  The output is what the first convolution wants to read: the 6x6/s2 stem equals a 3x3/s1
// convolution over the 2x2 space-to-depth image (channel (py*2+px)*3+c); K1 also packs the three COLUMN taps
// into the channel dimension, [H/2][W/2][48] fp16 (3 x (12 used + 4 zero)), so that the stem becomes a
// 3x1 convolution whose operand rows are full 128-byte swizzle rows (measured: 9 mostly zero-filled 64-channel
// chunks cost as much shared-memory fill as 9 real ones; 3 real chunks cost a third).
//
// HBM traffic per 1920x1080 -> 768x1280 image: 6.22 MB read (each source byte once, staged through shared
// memory with 128-bit loads) + 23.6 MB written.
#include <algorithm>
#include <cmath>

#include "common.h"

namespace mdb {

__constant__ __half c_norm_lut[256];  // half(float(v) / 255.0f)

void preprocess_init_constants() {
  __half lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = __float2half_rn(static_cast<float>(v) / 255.0f);
  MDB_CUDA(cudaMemcpyToSymbol(c_norm_lut, lut, sizeof(lut)));
}

LetterboxGeom letterbox_geometry(int h, int w, int img_size, int stride, bool autopad, bool scaleup) {
  // pytorch_detector.py:434-452 (the reference's own restatement of yolov5 letterbox geometry).
  // Python's round() is round-half-to-even on doubles == nearbyint in the default rounding mode.
  LetterboxGeom g{};
  g.src_h = h;
  g.src_w = w;
  double r = std::min(static_cast<double>(img_size) / h, static_cast<double>(img_size) / w);
  if (!scaleup) r = std::min(r, 1.0);
  g.ratio = r;
  g.new_w = static_cast<int>(std::nearbyint(w * r));
  g.new_h = static_cast<int>(std::nearbyint(h * r));
  double dw = img_size - g.new_w, dh = img_size - g.new_h;
  if (autopad) {
    dw = std::fmod(dw, stride);
    dh = std::fmod(dh, stride);
    if (dw < 0) dw += stride;  // np.mod semantics
    if (dh < 0) dh += stride;
  }
  dw /= 2;
  dh /= 2;
  g.dw = dw;
  g.dh = dh;
  g.top = static_cast<int>(std::nearbyint(dh - 0.1));
  g.bottom = static_cast<int>(std::nearbyint(dh + 0.1));
  g.left = static_cast<int>(std::nearbyint(dw - 0.1));
  g.right = static_cast<int>(std::nearbyint(dw + 0.1));
  g.out_h = g.new_h + g.top + g.bottom;
  g.out_w = g.new_w + g.left + g.right;
  return g;
}

LetterboxGeom letterbox_geometry_modern(int h, int w, int img_size, int stride, bool use_ceil) {
  // pytorch_detector.py:1036-1101 + yolov5 letterbox(auto=False, scaleup=False) on the pre-resized image
  LetterboxGeom g{};
  g.src_h = h;
  g.src_w = w;
  const double r = static_cast<double>(img_size) / std::max(h, w);
  g.new_w = w;
  g.new_h = h;
  g.interp = 0;
  if (r != 1.0) {
    g.interp = r > 1.0 ? 0 : 1;
    g.new_w = use_ceil ? static_cast<int>(std::ceil(w * r)) : static_cast<int>(w * r);
    g.new_h = use_ceil ? static_cast<int>(std::ceil(h * r)) : static_cast<int>(h * r);
  }
  // target_shape = ceil(normalized_shape * image_size / stride + 0.5) * stride, normalized by max(H, W, 3) of the resized image
  const double maxd = std::max(std::max(g.new_h, g.new_w), 3);
  const int th = static_cast<int>(std::ceil((g.new_h / maxd) * img_size / stride + 0.5)) * stride;
  const int tw = static_cast<int>(std::ceil((g.new_w / maxd) * img_size / stride + 0.5)) * stride;
  // letterbox(): r = min(th / H, tw / W) clipped to 1 (scaleup=False); th >= H and tw >= W, so no second resize
  g.ratio = 1.0;
  double dw = tw - g.new_w, dh = th - g.new_h;
  dw /= 2;
  dh /= 2;
  g.dw = dw;
  g.dh = dh;
  g.top = static_cast<int>(std::nearbyint(dh - 0.1));
  g.bottom = static_cast<int>(std::nearbyint(dh + 0.1));
  g.left = static_cast<int>(std::nearbyint(dw - 0.1));
  g.right = static_cast<int>(std::nearbyint(dw + 0.1));
  g.out_h = g.new_h + g.top + g.bottom;
  g.out_w = g.new_w + g.left + g.right;
  // scale_coords(img1_shape, coords, scaling_shape, ratio_pad=((H'/H, W'/W), (dw, dh))): gain = ratio_pad[0][0], pad = ratio_pad[1]
  g.sc_gain = static_cast<double>(g.new_h) / h;
  g.sc_pad_x = dw;
  g.sc_pad_y = dh;
  return g;
}

void area_tables(int ssize, int dsize, std::vector<int32_t>& first, std::vector<int32_t>& si, std::vector<float>& alpha) {
  // cv::computeResizeAreaTab (imgproc/src/resize.cpp); scale = 1 / (dsize / ssize) as cv::resize computes it
  const double inv_scale = static_cast<double>(dsize) / ssize;
  const double scale = 1.0 / inv_scale;
  first.assign(dsize + 1, 0);
  si.clear();
  alpha.clear();
  for (int dx = 0; dx < dsize; ++dx) {
    first[dx] = static_cast<int32_t>(si.size());
    const double fsx1 = dx * scale;
    const double fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = static_cast<int>(std::ceil(fsx1)), sx2 = static_cast<int>(std::floor(fsx2));
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) { si.push_back(sx1 - 1); alpha.push_back(static_cast<float>((sx1 - fsx1) / cell)); }
    for (int sx = sx1; sx < sx2; ++sx) { si.push_back(sx); alpha.push_back(static_cast<float>(1.0 / cell)); }
    if (fsx2 - sx2 > 1e-3) { si.push_back(sx2); alpha.push_back(static_cast<float>(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)); }
  }
  first[dsize] = static_cast<int32_t>(si.size());
}

void linear_tables(int ssize, int dsize, bool vertical, std::vector<int32_t>& ofs, std::vector<int16_t>& coef) {
  // cv::resize, INTER_LINEAR, 8-bit: imgproc/src/resize.cpp (coefficient set-up before resizeGeneric_).
  ofs.resize(dsize);
  coef.resize(2 * static_cast<size_t>(dsize));
  const double inv_scale = static_cast<double>(dsize) / ssize;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsize; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(std::floor(f));
    f -= static_cast<float>(s);
    if (!vertical) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    }
    ofs[d] = s;
    coef[2 * d] = static_cast<int16_t>(std::lrintf((1.f - f) * 2048.f));
    coef[2 * d + 1] = static_cast<int16_t>(std::lrintf(f * 2048.f));
  }
}

constexpr int kPreThreads = 256;
constexpr int kTileW2 = 64;  // s2d columns per CTA  (128 output columns)
constexpr int kTileH2 = 8;   // s2d rows per CTA     (16 output rows)
constexpr int kHaloW2 = kTileW2 + 2;  // one extra s2d column on each side: the stem's column taps
constexpr int kStemChannels = 48;     // 3 taps x 16 (12 used); the conv's TMA zero-fills channels 48..63

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Output layout (what the stem convolution reads, see pack_weights_stem_s2d): per s2d pixel (y, x) 48 fp16 channels,
//   [t*16 + (py*2+px)*3 + c]  = normalised letterboxed pixel (2y+py, 2(x+t-1)+px), channel c, for t = 0,1,2
// i.e. the three column taps of the stem packed into the channel dimension (zero outside the image = the conv's
// zero padding; channels 12..15 of every tap are zero; the conv reads 64-channel boxes, 48..63 zero-filled by TMA).
// STAGED: the source window of the tile is first copied to shared memory with 16-byte loads.
template <bool STAGED>
__global__ void __launch_bounds__(kPreThreads) letterbox_s2d_kernel(const PreImage* __restrict__ imgs, int out_h,
                                                                   int out_w, __half* __restrict__ dst,
                                                                   int smem_row_bytes, int smem_rows) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  // u8 -> half(v / 255) table in shared memory: the lanes of a warp look up DIFFERENT entries, which the constant cache
  // serialises (up to 32 replays per lookup, 12 lookups per s2d pixel); shared memory serves them with few conflicts
  __shared__ __half s_lut[256];
  if (threadIdx.x < 256) s_lut[threadIdx.x] = c_norm_lut[threadIdx.x];
  if (!STAGED) __syncthreads();
  // [kTileH2][kHaloW2] entries of 16 halves (12 used): the letterboxed, normalised s2d tile with column halo
  uint4* s_tile = reinterpret_cast<uint4*>(s_dyn);
  uint8_t* s_src = s_dyn + static_cast<size_t>(kTileH2) * kHaloW2 * 32;
  const PreImage im = imgs[blockIdx.z];
  const int W2 = out_w >> 1, H2 = out_h >> 1;
  const int sx2_0 = blockIdx.x * kTileW2, sy2_0 = blockIdx.y * kTileH2;

  // output-pixel window (with the column halo) of this tile that lies inside the resized image
  const int ox_lo = max(2 * (sx2_0 - 1), im.left), ox_hi = min(2 * (sx2_0 + kTileW2 + 1), im.left + im.new_w) - 1;
  const int oy_lo = max(2 * sy2_0, im.top), oy_hi = min(2 * (sy2_0 + kTileH2), im.top + im.new_h) - 1;
  const bool any_src = (ox_lo <= ox_hi) && (oy_lo <= oy_hi);

  int row0 = 0, byte0 = 0;  // source window origin
  if (STAGED) {
    if (any_src) {
      int row1, col0, col1;
      if (im.mode == 0) {
        row0 = clampi(im.yofs[oy_lo - im.top], 0, im.src_h - 1);
        row1 = clampi(im.yofs[oy_hi - im.top] + 1, 0, im.src_h - 1);
        col0 = im.xofs[ox_lo - im.left];
        col1 = min(im.xofs[ox_hi - im.left] + 1, im.src_w - 1);
      } else if (im.mode == 1) {
        row0 = im.ya_si[im.ya_first[oy_lo - im.top]];
        row1 = im.ya_si[im.ya_first[oy_hi - im.top + 1] - 1];
        col0 = im.xa_si[im.xa_first[ox_lo - im.left]];
        col1 = im.xa_si[im.xa_first[ox_hi - im.left + 1] - 1];
      } else {
        row0 = (oy_lo - im.top) * im.iscale_y;
        row1 = min((oy_hi - im.top + 1) * im.iscale_y - 1, im.src_h - 1);
        col0 = (ox_lo - im.left) * im.iscale_x;
        col1 = min((ox_hi - im.left + 1) * im.iscale_x - 1, im.src_w - 1);
      }
      byte0 = col0 * 3;
      const int nbytes = (col1 + 1) * 3 - byte0;
      const int nrows = row1 - row0 + 1;
      // each row: 16-byte vectors starting at the aligned-down address of its first needed byte
      const int vec_per_row = smem_row_bytes >> 4;
      for (int i = threadIdx.x; i < nrows * vec_per_row; i += kPreThreads) {
        const int r = i / vec_per_row, v = i - r * vec_per_row;
        const uint8_t* p = im.src + static_cast<long long>(row0 + r) * im.src_stride + byte0;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~static_cast<uintptr_t>(15);
        const int shift = static_cast<int>(reinterpret_cast<uintptr_t>(p) - a);
        if (v * 16 < shift + nbytes) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(a) + v);
          *reinterpret_cast<uint4*>(s_src + static_cast<size_t>(r) * smem_row_bytes + v * 16) = q;
        }
      }
      (void)smem_rows;
    }
    __syncthreads();
  }

  // start of source row sy (byte 0 of the row) as an offset into s_src / a pointer into global memory
  auto row_ptr = [&](int sy) -> const uint8_t* {
    const uint8_t* g = im.src + static_cast<long long>(sy) * im.src_stride;
    if (STAGED) {
      const int shift = static_cast<int>(reinterpret_cast<uintptr_t>(g + byte0) & 15);
      return s_src + static_cast<size_t>(sy - row0) * smem_row_bytes + shift - byte0;
    }
    return g;
  };

  // ---- phase 1: letterbox (OpenCV-exact bilinear or pad) every s2d pixel of the tile + halo columns ----
  for (int t = threadIdx.x; t < kHaloW2 * kTileH2; t += kPreThreads) {
    const int lx = t % kHaloW2, ly = t / kHaloW2;
    const int sx2 = sx2_0 + lx - 1, sy2 = sy2_0 + ly;
    __align__(16) __half px[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) px[i] = __ushort_as_half(0);
    if (sx2 >= 0 && sx2 < W2 && sy2 < H2) {
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const int oy = 2 * sy2 + py;
        const int dy = oy - im.top;
        const bool in_y = dy >= 0 && dy < im.new_h;
        int b0 = 0, b1 = 0;
        const uint8_t* r0 = nullptr;
        const uint8_t* r1 = nullptr;
        if (in_y && im.mode == 0) {
          const int s = im.yofs[dy];
          r0 = row_ptr(clampi(s, 0, im.src_h - 1));
          r1 = row_ptr(clampi(s + 1, 0, im.src_h - 1));
          b0 = im.ycoef[2 * dy];
          b1 = im.ycoef[2 * dy + 1];
        }
#pragma unroll
        for (int pxx = 0; pxx < 2; ++pxx) {
          const int ox = 2 * sx2 + pxx;
          const int dx = ox - im.left;
          int v[3] = {114, 114, 114};
          if (im.mode != 0) {
            if (dy >= 0 && dy < im.new_h && dx >= 0 && dx < im.new_w) {
              if (im.mode == 1) {
                // cv::ResizeArea_: buf[dx] = sum_k S[sx_k] * alpha_k per source row (float, in table order), then
                // sum = beta * buf for the first row of the group and sum += beta * buf for the others; cvRound at the end
                float acc[3] = {0.f, 0.f, 0.f};
                const int y0 = im.ya_first[dy], y1 = im.ya_first[dy + 1];
                const int x0 = im.xa_first[dx], x1 = im.xa_first[dx + 1];
                for (int ky = y0; ky < y1; ++ky) {
                  const uint8_t* rr = row_ptr(im.ya_si[ky]);
                  const float beta = im.ya_alpha[ky];
                  float b[3] = {0.f, 0.f, 0.f};
                  for (int kx = x0; kx < x1; ++kx) {
                    const int o = im.xa_si[kx] * 3;
                    const float a = im.xa_alpha[kx];
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                      b[c] = __fadd_rn(b[c], __fmul_rn(static_cast<float>(STAGED ? rr[o + c] : __ldg(rr + o + c)), a));
                  }
#pragma unroll
                  for (int c = 0; c < 3; ++c)
                    acc[c] = (ky == y0) ? __fmul_rn(beta, b[c]) : __fadd_rn(acc[c], __fmul_rn(beta, b[c]));
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = min(255, max(0, __float2int_rn(acc[c])));
              } else {
                // cv::ResizeAreaFast_: integer block sums; 2x2: (s + 2) >> 2, otherwise cvRound(s * (1.f / area))
                int sum[3] = {0, 0, 0};
                for (int ky = 0; ky < im.iscale_y; ++ky) {
                  const uint8_t* rr = row_ptr(dy * im.iscale_y + ky);
                  for (int kx = 0; kx < im.iscale_x; ++kx) {
                    const int o = (dx * im.iscale_x + kx) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) sum[c] += STAGED ? rr[o + c] : __ldg(rr + o + c);
                  }
                }
                if (im.iscale_x == 2 && im.iscale_y == 2) {
#pragma unroll
                  for (int c = 0; c < 3; ++c) v[c] = (sum[c] + 2) >> 2;
                } else {
                  const float sc = __fdiv_rn(1.0f, static_cast<float>(im.iscale_x * im.iscale_y));
#pragma unroll
                  for (int c = 0; c < 3; ++c) v[c] = min(255, max(0, __float2int_rn(__fmul_rn(static_cast<float>(sum[c]), sc))));
                }
              }
            }
          } else if (in_y && dx >= 0 && dx < im.new_w) {
            const int sx0 = im.xofs[dx];
            const int o0 = sx0 * 3, o1 = min(sx0 + 1, im.src_w - 1) * 3;
            const int a0 = im.xcoef[2 * dx], a1 = im.xcoef[2 * dx + 1];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int h0 = (STAGED ? r0[o0 + c] : __ldg(r0 + o0 + c)) * a0 + (STAGED ? r0[o1 + c] : __ldg(r0 + o1 + c)) * a1;
              const int h1 = (STAGED ? r1[o0 + c] : __ldg(r1 + o0 + c)) * a0 + (STAGED ? r1[o1 + c] : __ldg(r1 + o1 + c)) * a1;
              v[c] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) px[(py * 2 + pxx) * 3 + c] = s_lut[v[c] & 255];
        }
      }
    }
    s_tile[t * 2] = reinterpret_cast<const uint4*>(px)[0];
    s_tile[t * 2 + 1] = reinterpret_cast<const uint4*>(px)[1];
  }
  __syncthreads();

  // ---- phase 2: one packed 96-byte row per s2d pixel (taps x-1, x, x+1).  Consecutive lanes write consecutive
  // 16-byte pieces, so a warp store covers 512 contiguous bytes.
  constexpr int kPieces = 6;
  for (int i = threadIdx.x; i < kTileW2 * kTileH2 * kPieces; i += kPreThreads) {
    const int piece = i % kPieces, pxl = i / kPieces;
    const int lx = pxl % kTileW2, ly = pxl / kTileW2;
    const int sx2 = sx2_0 + lx, sy2 = sy2_0 + ly;
    if (sx2 >= W2 || sy2 >= H2) continue;
    // piece p = half (p & 1) of the entry of column x - 1 + p / 2
    const uint4 v = s_tile[(ly * kHaloW2 + lx + (piece >> 1)) * 2 + (piece & 1)];
    reinterpret_cast<uint4*>(dst + ((static_cast<size_t>(blockIdx.z) * H2 + sy2) * W2 + sx2) * kStemChannels)[piece] = v;
  }
}

void preprocess_launch(const PreImage* d_imgs, int n, int out_h, int out_w, __half* dst, float max_scale,
                       cudaStream_t st) {
  if (n <= 0) return;
  const int W2 = out_w / 2, H2 = out_h / 2;
  dim3 grid((W2 + kTileW2 - 1) / kTileW2, (H2 + kTileH2 - 1) / kTileH2, n);
  // source window of one tile: (2*tile+1)*scale + 2 pixels each way, plus alignment slack
  const int src_cols = static_cast<int>(std::ceil((2 * kHaloW2 + 1) * max_scale)) + 3;
  const int src_rows = static_cast<int>(std::ceil((2 * kTileH2 + 1) * max_scale)) + 3;
  const int row_bytes = ((src_cols * 3 + 15 + 15) / 16) * 16;
  const size_t tile_bytes = static_cast<size_t>(kTileH2) * kHaloW2 * 32;
  const size_t smem = tile_bytes + static_cast<size_t>(row_bytes) * src_rows;
  if (smem <= 200 * 1024) {  // opted in per device by preprocess_init_device()
    letterbox_s2d_kernel<true><<<grid, kPreThreads, smem, st>>>(d_imgs, out_h, out_w, dst, row_bytes, src_rows);
  } else {
    // extreme down-scales (source window larger than shared memory): read through L1 instead
    letterbox_s2d_kernel<false><<<grid, kPreThreads, tile_bytes, st>>>(d_imgs, out_h, out_w, dst, 0, 0);
  }
  MDB_CUDA(cudaGetLastError());
}

// per-device setup (mdb_create, after cudaSetDevice): cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute
void preprocess_init_device() {
  preprocess_init_constants();
  MDB_CUDA(cudaFuncSetAttribute(letterbox_s2d_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
}

}  // namespace mdb
