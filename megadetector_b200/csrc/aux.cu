// K3 nearest-2x upsample into a concat buffer, K4 SPPF max-pools, K5 Detect decode.
// All three are pure data movement / elementwise work on NHWC tensors: HBM-bound, 16-byte vector accesses,
// one thread per (pixel, 8-channel group).
#include <algorithm>
#include <cfloat>

#include "common.h"
#include "nms_cand.cuh"

namespace mdb {

// ---------------------------------------------------------------------------------------------
// K3: nn.Upsample(scale_factor=2, mode='nearest') (yolov5x6.yaml idx 13/17/21), written
// straight into channels [c_off, c_off+C) of the following Concat's buffer.
// ---------------------------------------------------------------------------------------------
// One thread per SOURCE pixel and 8-channel group: one 16-byte load, four 16-byte stores (the 2x2 destination block);
// grid.y walks the source rows so that no 64-bit division is needed (the first version spent its time on index arithmetic:
// 73 % issue-slot utilisation at 44 % of the HBM rate, profiles/r02_ncu_upsample2x_key_metrics.csv).
__global__ void upsample2x_kernel(const __half* __restrict__ src, long long src_px, int src_coff,
                                  __half* __restrict__ dst, long long dst_px, int dst_coff, int Hs, int Ws, int C8) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Ws * C8) return;
  const int x = t / C8, c8 = t - x * C8;
  const long long row = blockIdx.y;  // n * Hs + y
  const long long n = row / Hs;
  const int y = static_cast<int>(row - n * Hs);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (row * Ws + x) * src_px + src_coff) + c8);
  const int Wd = 2 * Ws;
  __half* d0 = dst + ((n * 2 * Hs + 2 * y) * Wd + 2 * x) * dst_px + dst_coff;
  __half* d1 = d0 + static_cast<long long>(Wd) * dst_px;
  reinterpret_cast<uint4*>(d0)[c8] = v;
  reinterpret_cast<uint4*>(d0 + dst_px)[c8] = v;
  reinterpret_cast<uint4*>(d1)[c8] = v;
  reinterpret_cast<uint4*>(d1 + dst_px)[c8] = v;
}

void upsample2x_launch(const View& src, const View& dst, int n_img, cudaStream_t st) {
  MDB_REQUIRE(src.C == dst.C && src.C % 8 == 0 && dst.H == 2 * src.H && dst.W == 2 * src.W, "upsample view mismatch");
  const int C8 = src.C / 8;
  if (n_img <= 0 || src.H <= 0 || src.W <= 0) return;
  MDB_REQUIRE(static_cast<long long>(n_img) * src.H <= 65535, "too many rows for one upsample launch");
  const int threads = 256;
  dim3 grid(static_cast<unsigned>((src.W * C8 + threads - 1) / threads), static_cast<unsigned>(n_img * src.H));
  upsample2x_kernel<<<grid, threads, 0, st>>>(static_cast<const __half*>(src.base), src.Ctot, src.c_off,
                                              static_cast<__half*>(dst.base), dst.Ctot, dst.c_off, src.H, src.W, C8);
  MDB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// K4: SPPF (yolov5 models/common.py SPPF, k=5): p1 = maxpool5(s1), p2 = maxpool5(p1), p3 = maxpool5(p2),
// all stride 1 / pad 2 (-inf padding), i.e. 5x5, 9x9 and 13x13 window maxima of s1.  One pass.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void hmax8(uint4& a, const uint4& b) {
  __half2* pa = reinterpret_cast<__half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) pa[i] = __hmax2(pa[i], pb[i]);
}

__global__ void sppf_pool_kernel(const __half* __restrict__ s1, long long px, int c_s1, __half* __restrict__ out,
                                 int c_p1, int c_p2, int c_p3, int H, int W, int C8, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % C8);
  long long p = i / C8;
  const int x = static_cast<int>(p % W);
  p /= W;
  const int y = static_cast<int>(p % H);
  const long long n = p / H;
  const __half2 ninf = __float2half2_rn(-INFINITY);
  uint4 m5, m9, m13;
  *reinterpret_cast<__half2*>(&m5.x) = ninf; *reinterpret_cast<__half2*>(&m5.y) = ninf;
  *reinterpret_cast<__half2*>(&m5.z) = ninf; *reinterpret_cast<__half2*>(&m5.w) = ninf;
  m9 = m5;
  m13 = m5;
  for (int dy = -6; dy <= 6; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -6; dx <= 6; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(s1 + ((n * H + yy) * W + xx) * px + c_s1) + c8);
      hmax8(m13, v);
      const int ady = dy < 0 ? -dy : dy, adx = dx < 0 ? -dx : dx;
      if (ady <= 4 && adx <= 4) hmax8(m9, v);
      if (ady <= 2 && adx <= 2) hmax8(m5, v);
    }
  }
  __half* o = out + ((n * H + y) * W + x) * px;
  reinterpret_cast<uint4*>(o + c_p1)[c8] = m5;
  reinterpret_cast<uint4*>(o + c_p2)[c8] = m9;
  reinterpret_cast<uint4*>(o + c_p3)[c8] = m13;
}

// Separable form for feature maps that fit shared memory (always, at MegaDetector's sizes: 12 x 20 at 768 x 1280): one CTA
// owns the H x W plane of one image and one 8-channel group, and applies the 5-wide row maximum and the 5-high column maximum
// three times (maxpool5 o maxpool5 o maxpool5, exactly the reference's chain) -- 30 comparisons per output instead of the
// 169 loads of the direct 13 x 13 window (which ran at 1.7 % of the HBM rate, profiles/r02_ncu_sppf_pool_key_metrics.csv).
__global__ void sppf_plane_kernel(const __half* __restrict__ s1, long long px, int c_s1, __half* __restrict__ out, int c_p1,
                                  int c_p2, int c_p3, int H, int W, int C8) {
  extern __shared__ uint4 s_plane[];
  uint4* A = s_plane;
  uint4* B = s_plane + H * W;
  const int c8 = blockIdx.x % C8;
  const long long n = blockIdx.x / C8;
  const int HW = H * W;
  const __half* base = s1 + n * HW * px;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) A[i] = __ldg(reinterpret_cast<const uint4*>(base + i * px + c_s1) + c8);
  __syncthreads();
  const int offs[3] = {c_p1, c_p2, c_p3};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int y = i / W, x = i - y * W;
      uint4 m = A[i];
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (dx != 0 && xx >= 0 && xx < W) hmax8(m, A[y * W + xx]);
      }
      B[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int y = i / W, x = i - y * W;
      uint4 m = B[i];
      for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (dy != 0 && yy >= 0 && yy < H) hmax8(m, B[yy * W + x]);
      }
      A[i] = m;
      reinterpret_cast<uint4*>(out + (n * HW + i) * px + offs[pass])[c8] = m;
    }
    __syncthreads();
  }
}

void sppf_pool_launch(const View& s1, const View& p1, const View& p2, const View& p3, int n_img, cudaStream_t st) {
  MDB_REQUIRE(s1.base == p1.base && s1.base == p2.base && s1.base == p3.base, "SPPF views must share the concat buffer");
  MDB_REQUIRE(s1.C % 8 == 0, "SPPF channels must be a multiple of 8");
  const int C8 = s1.C / 8;
  {
    const size_t smem = static_cast<size_t>(s1.H) * s1.W * 2 * sizeof(uint4);
    if (n_img > 0 && smem <= 48 * 1024) {
      const int threads = std::min(256, std::max(32, ((s1.H * s1.W + 31) / 32) * 32));
      sppf_plane_kernel<<<static_cast<unsigned>(n_img * C8), threads, smem, st>>>(
          static_cast<const __half*>(s1.base), s1.Ctot, s1.c_off, static_cast<__half*>(s1.base), p1.c_off, p2.c_off, p3.c_off,
          s1.H, s1.W, C8);
      MDB_CUDA(cudaGetLastError());
      return;
    }
  }
  const long long total = static_cast<long long>(n_img) * s1.H * s1.W * C8;
  if (total == 0) return;
  const int threads = 128;
  sppf_pool_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, st>>>(
      static_cast<const __half*>(s1.base), s1.Ctot, s1.c_off, static_cast<__half*>(s1.base), p1.c_off, p2.c_off,
      p3.c_off, s1.H, s1.W, C8, total);
  MDB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// K5: Detect inference decode (yolov5 models/yolo.py Detect.forward, v7.0 era = ultralytics-yolov5 0.1.1;
// restated in oracle/yolov5.py::decode_level):   y = sigmoid(raw)
//     xy = (y*2 + (grid - 0.5)) * stride        wh = (y*2)^2 * anchor_px        rest = y
// row of pred = level_off + (a*ny + iy)*nx + ix.  Every product/sum is rounded separately (no FMA
// contraction) so the op order is the reference's.
// ---------------------------------------------------------------------------------------------
struct DecodeArgs {
  DecodeLevel lv[4];
};

__device__ __forceinline__ float sigmoid_acc(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// With `keys != nullptr` the kernel also runs K6 on the row it just produced (confidence threshold, best class,
// candidate key append; see nms.cu): the prediction tensor is thresholded at HBM rate by the thread that wrote it.
// TTA (yolov5 _forward_augment): a pass writes its rows at `out_base` of a longer per-image tensor of `out_rows` rows, with
// _descale_pred applied: xywh /= scale, and x = flip_w - x for the left-right flipped pass (flip_w == 0: no flip).
__global__ void decode_kernel(const DecodeArgs args, int no, int rows_per_img, float* __restrict__ pred,
                              long long total_rows, float thr, unsigned long long* __restrict__ keys,
                              int32_t* __restrict__ counts, int cap, int out_base, int out_rows, float scale, float flip_w) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool in = r < total_rows;
  int n = -1, row = 0;
  bool ok = false;
  float conf = 0.0f;
  if (in) {
    n = static_cast<int>(r / rows_per_img);
    row = static_cast<int>(r - static_cast<long long>(n) * rows_per_img);
    int l = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (row >= args.lv[i].row_off) l = i;
    const DecodeLevel& L = args.lv[l];
    const int local = row - L.row_off;
    const int ix = local % L.nx;
    const int iy = (local / L.nx) % L.ny;
    const int a = local / (L.nx * L.ny);
    const float* src = L.raw + ((static_cast<long long>(n) * L.ny + iy) * L.nx + ix) * L.c_stride + a * no;
    row += out_base;
    float* dst = pred + (static_cast<long long>(n) * out_rows + row) * no;
    const float s0 = sigmoid_acc(src[0]), s1 = sigmoid_acc(src[1]), s2 = sigmoid_acc(src[2]), s3 = sigmoid_acc(src[3]);
    float bx = __fmul_rn(__fadd_rn(__fmul_rn(s0, 2.0f), static_cast<float>(ix) - 0.5f), L.stride);
    float by = __fmul_rn(__fadd_rn(__fmul_rn(s1, 2.0f), static_cast<float>(iy) - 0.5f), L.stride);
    const float w2 = __fmul_rn(s2, 2.0f), h2 = __fmul_rn(s3, 2.0f);
    float bw = __fmul_rn(__fmul_rn(w2, w2), L.anchors[a][0]);
    float bh = __fmul_rn(__fmul_rn(h2, h2), L.anchors[a][1]);
    if (scale != 1.0f) {
      bx = __fdiv_rn(bx, scale); by = __fdiv_rn(by, scale); bw = __fdiv_rn(bw, scale); bh = __fdiv_rn(bh, scale);
    }
    if (flip_w != 0.0f) bx = __fsub_rn(flip_w, bx);
    dst[0] = bx; dst[1] = by; dst[2] = bw; dst[3] = bh;
    const float obj = sigmoid_acc(src[4]);
    dst[4] = obj;
    float best = -INFINITY;
    for (int c = 5; c < no; ++c) {
      const float p = sigmoid_acc(src[c]);
      dst[c] = p;
      // reference nms(): x[:, 5:] *= x[:, 4:5]; conf = max over classes (first maximum), pytorch_detector.py:560-575
      const float cc = __fmul_rn(p, obj);
      if (cc > best) best = cc;
    }
    ok = keys != nullptr && obj > thr && best > thr;
    conf = best;
  }
  if (keys != nullptr) cand_append(ok, n, row, conf, keys, counts, cap);
}

void decode_launch(const DecodeLevel* lv4, int n_img, int no, int rows_per_img, float* pred, cudaStream_t st, float thr,
                   unsigned long long* cand_keys, int32_t* cand_counts, int cand_cap, const DecodePass* pass) {
  DecodeArgs args;
  for (int i = 0; i < 4; ++i) args.lv[i] = lv4[i];
  const long long total = static_cast<long long>(n_img) * rows_per_img;
  if (total == 0) return;
  const bool first = pass == nullptr || pass->out_base == 0;
  if (cand_keys != nullptr && first) MDB_CUDA(cudaMemsetAsync(cand_counts, 0, sizeof(int32_t) * n_img, st));
  const int threads = 256;
  decode_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, st>>>(
      args, no, rows_per_img, pred, total, thr, cand_keys, cand_counts, cand_cap, pass ? pass->out_base : 0,
      pass ? pass->out_rows : rows_per_img, pass ? pass->scale : 1.0f, pass ? pass->flip_w : 0.0f);
  MDB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// TTA input: yolov5 utils/torch_utils.py scale_img(x.flip(3) if flip else x, ratio, gs=64) on the normalised image:
// F.interpolate(size=(int(H*r), int(W*r)), mode='bilinear', align_corners=False), then pad right/bottom with 0.447 to the
// next multiple of 64.  Source = the stem-input tensor K1 wrote (its centre tap holds every pixel's fp16(v / 255): v is
// recovered exactly and the interpolation runs on float(v) / 255 in fp32 like the reference's); destination = the stem-input
// tensor of the plan of the scaled shape, in the same tap-packed space-to-depth layout.
// ---------------------------------------------------------------------------------------------
__global__ void tta_scale_s2d_kernel(const __half* __restrict__ src, int H, int W, __half* __restrict__ dst, int Hd, int Wd,
                                     int hs, int ws, float sc_h, float sc_w, int flip, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W2d = Wd >> 1, H2d = Hd >> 1, W2 = W >> 1;
  const int x2 = static_cast<int>(i % W2d);
  const int y2 = static_cast<int>((i / W2d) % H2d);
  const long long n = i / (static_cast<long long>(W2d) * H2d);
  const __half* S = src + n * (H >> 1) * W2 * 48;
  auto pix = [&](int sy, int sx, int c) -> float {  // exact float(v) / 255 of source pixel (sy, sx), channel c
    if (flip) sx = W - 1 - sx;
    const __half h = S[(static_cast<long long>(sy >> 1) * W2 + (sx >> 1)) * 48 + 16 + ((sy & 1) * 2 + (sx & 1)) * 3 + c];
    const float v = rintf(__half2float(h) * 255.0f);
    return __fdiv_rn(v, 255.0f);
  };
  __half out[12];
#pragma unroll
  for (int py = 0; py < 2; ++py) {
    const int Y = 2 * y2 + py;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int X = 2 * x2 + px;
      float v[3] = {0.447f, 0.447f, 0.447f};
      if (Y < hs && X < ws) {
        // PyTorch upsample_bilinear2d, align_corners=False: src = scale * (dst + 0.5) - 0.5 clamped at 0
        float fy = sc_h * (static_cast<float>(Y) + 0.5f) - 0.5f;
        float fx = sc_w * (static_cast<float>(X) + 0.5f) - 0.5f;
        fy = fy < 0.f ? 0.f : fy;
        fx = fx < 0.f ? 0.f : fx;
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[c] = ly0 * (lx0 * pix(y0, x0, c) + lx1 * pix(y0, x1, c)) + ly1 * (lx0 * pix(y1, x0, c) + lx1 * pix(y1, x1, c));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) out[(py * 2 + px) * 3 + c] = __float2half_rn(v[c]);
    }
  }
  // this pixel's 2x2 block is tap 1 of itself, tap 0 of its right neighbour and tap 2 of its left neighbour
  __half* D = dst + (n * H2d + y2) * static_cast<long long>(W2d) * 48;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int xd = x2 + 1 - t;
    if (xd < 0 || xd >= W2d) continue;
    __half* d = D + static_cast<long long>(xd) * 48 + t * 16;
#pragma unroll
    for (int k = 0; k < 12; ++k) d[k] = out[k];
  }
}

void tta_scale_launch(const __half* src, int H, int W, __half* dst, int Hd, int Wd, double ratio, bool flip, int n_img,
                      cudaStream_t st) {
  const int hs = static_cast<int>(H * ratio), ws = static_cast<int>(W * ratio);
  MDB_REQUIRE(hs >= 1 && ws >= 1 && hs <= Hd && ws <= Wd, "TTA scaled size does not fit its plan");
  const long long total = static_cast<long long>(n_img) * (Hd >> 1) * (Wd >> 1);
  if (total == 0) return;
  const float sc_h = static_cast<float>(H) / static_cast<float>(hs), sc_w = static_cast<float>(W) / static_cast<float>(ws);
  tta_scale_s2d_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, H, W, dst, Hd, Wd, hs, ws, sc_h, sc_w,
                                                                                   flip ? 1 : 0, total);
  MDB_CUDA(cudaGetLastError());
}

}  // namespace mdb
