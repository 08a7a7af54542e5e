// K6-K8: confidence threshold + best-class selection + per-class greedy NMS + top-k + box rescale.
//
// Replaces the reference's `nms()` (megadetector/detection/pytorch_detector.py:502-610, a loop of ATen
// kernels, torchvision.ops.nms and >= 3 host syncs per image) and `scale_coords(...).round()`
// (pytorch_detector.py:1379).  One CTA of 1024 threads owns one image and never talks to the host:
//
//   K6  ordered compaction   anchors in index order: obj > thr, conf = p_c*obj, first-max class, conf > thr
//                            (warp ballot + block scan; candidates stay in anchor order)
//   K7a stable LSD radix sort of the candidates by descending confidence (4 x 8 bits over the float bits;
//       warp match + per-warp digit counters); stability == torchvision's "equal score: lower index first"
//   K7b greedy suppression over the globally sorted list with class-aware IoU tests.  Because a box can only
//       be suppressed by a higher-ranked box of its own class, walking the merged list and stopping after
//       max_det survivors yields exactly the reference's "per-class NMS, concat, sort by conf, keep 300".
//   K8  scale_coords + clip + round-half-even on the survivors, written with their anchor indices.
//
// Bit-exactness: every float op that torch/torchvision perform is issued with an explicit round-to-nearest
// intrinsic, so nvcc cannot contract a*b+c into an FMA; thresholds are compared in the reference's types
// (conf: float32(thr); IoU: float promoted to double vs the double threshold == float > floor_f32(thr)).
#include <cmath>
#include <cstring>

#include "common.h"

namespace mdb {

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;

float float_floor_of_double(double t) {
  float f = static_cast<float>(t);
  if (static_cast<double>(f) > t) f = std::nextafterf(f, -INFINITY);
  return f;
}

size_t nms_scratch_bytes(int n_img, int anchors) { return static_cast<size_t>(n_img) * anchors * 16; }

void nms_bind_scratch(NmsScratch* s, void* mem, int n_img, int anchors) {
  char* p = static_cast<char*>(mem);
  const size_t one = static_cast<size_t>(n_img) * anchors * 4;
  s->key[0] = reinterpret_cast<uint32_t*>(p);
  s->key[1] = reinterpret_cast<uint32_t*>(p + one);
  s->idx[0] = reinterpret_cast<int32_t*>(p + 2 * one);
  s->idx[1] = reinterpret_cast<int32_t*>(p + 3 * one);
  s->cap = anchors;
}

struct Cand {
  float x1, y1, x2, y2, area, conf;
  int cls;
};

// reference: box = xc -/+ w/2 (float32), class_conf = p*obj, first maximum
__device__ __forceinline__ bool eval_anchor(const float* __restrict__ row, int no, float thr, Cand& c) {
  float v[8];
  if (no == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(row));
    const float4 b = __ldg(reinterpret_cast<const float4*>(row) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = __ldg(row + i);
  }
  const float obj = v[4];
  if (!(obj > thr)) return false;
  float best = -INFINITY;
  int bi = 0;
  if (no == 8) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float cc = __fmul_rn(v[5 + k], obj);
      if (cc > best) { best = cc; bi = k; }
    }
  } else {
    for (int k = 0; k < no - 5; ++k) {
      const float cc = __fmul_rn(__ldg(row + 5 + k), obj);
      if (cc > best) { best = cc; bi = k; }
    }
  }
  if (!(best > thr)) return false;
  const float hw = __fdiv_rn(v[2], 2.0f), hh = __fdiv_rn(v[3], 2.0f);
  c.x1 = __fsub_rn(v[0], hw);
  c.y1 = __fsub_rn(v[1], hh);
  c.x2 = __fadd_rn(v[0], hw);
  c.y2 = __fadd_rn(v[1], hh);
  c.area = __fmul_rn(__fsub_rn(c.x2, c.x1), __fsub_rn(c.y2, c.y1));
  c.conf = best;
  c.cls = bi;
  return true;
}

// torchvision CPU nms_kernel_impl: ovr = inter / (iarea + areas[j] - inter); suppressed if ovr > thr
__device__ __forceinline__ bool iou_exceeds(const Cand& a, const Cand& b, float thr_floor) {
  const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
  const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
  return ovr > thr_floor;
}

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_image_kernel(const float* __restrict__ pred, int anchors, int no, float thr, float iou_floor, int max_det,
                 const NmsImageParams* __restrict__ params, NmsScratch sc, float* __restrict__ det_out,
                 int32_t* __restrict__ det_count, int32_t* __restrict__ kept_anchor) {
  __shared__ uint32_t s_warp_cnt[kNmsWarps];
  __shared__ uint32_t s_digit_off[256];
  __shared__ uint16_t s_warp_digit[kNmsWarps][256];
  __shared__ int s_n, s_kept, s_first;
  __shared__ Cand s_keep[kMaxDetCap];
  __shared__ int s_keep_anchor[kMaxDetCap];
  __shared__ int s_warp_first[kNmsWarps];

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = pred + static_cast<size_t>(img) * anchors * no;
  uint32_t* key0 = sc.key[0] + static_cast<size_t>(img) * sc.cap;
  uint32_t* key1 = sc.key[1] + static_cast<size_t>(img) * sc.cap;
  int32_t* idx0 = sc.idx[0] + static_cast<size_t>(img) * sc.cap;
  int32_t* idx1 = sc.idx[1] + static_cast<size_t>(img) * sc.cap;

  // ------------------------------- K6: ordered compaction -------------------------------
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int base = 0; base < anchors; base += kNmsThreads) {
    const int a = base + tid;
    Cand c;
    const bool ok = (a < anchors) && eval_anchor(P + static_cast<size_t>(a) * no, no, thr, c);
    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w2 = 0; w2 < kNmsWarps; ++w2) {
      const uint32_t cnt = s_warp_cnt[w2];
      if (w2 < warp) before += cnt;
      total += cnt;
    }
    const int n0 = s_n;
    if (ok) {
      const int pos = n0 + before + __popc(bal & ((1u << lane) - 1u));
      key0[pos] = ~__float_as_uint(c.conf);  // conf > 0: bit pattern is monotonic; inverted for descending order
      idx0[pos] = a;
    }
    __syncthreads();
    if (tid == 0) s_n = n0 + total;
    __syncthreads();
  }
  const int n = s_n;

  // ------------------------------- K7a: stable LSD radix sort -------------------------------
  uint32_t* kin = key0; uint32_t* kout = key1;
  int32_t* iin = idx0; int32_t* iout = idx1;
  for (int pass = 0; pass < 4 && n > 1; ++pass) {
    const int shift = pass * 8;
    if (tid < 256) s_digit_off[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kNmsThreads) atomicAdd(&s_digit_off[(kin[i] >> shift) & 255u], 1u);
    __syncthreads();
    // exclusive scan over 256 digit counts (warp 0, 8 per lane)
    if (warp == 0) {
      uint32_t loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = s_digit_off[lane * 8 + j]; sum += loc[j]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s_digit_off[lane * 8 + j] = run; run += loc[j]; }
    }
    __syncthreads();
    for (int base = 0; base < n; base += kNmsThreads) {
      const int i = base + tid;
      const bool valid = i < n;
      uint32_t k = 0; int32_t v = 0; uint32_t d = 0;
      if (valid) { k = kin[i]; v = iin[i]; d = (k >> shift) & 255u; }
      for (int j = lane; j < 256; j += 32) s_warp_digit[warp][j] = 0;
      __syncwarp();
      // peers in this warp with the same digit (invalid lanes use digit 256 -> never match a valid one)
      const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 256u);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      if (valid && rank == 0) s_warp_digit[warp][d] = static_cast<uint16_t>(__popc(peers));
      __syncthreads();
      uint32_t pos = 0;
      if (valid) {
        uint32_t before = 0;
        for (int w2 = 0; w2 < warp; ++w2) before += s_warp_digit[w2][d];
        pos = s_digit_off[d] + before + rank;
      }
      __syncthreads();
      if (tid < 256) {
        uint32_t tot = 0;
        for (int w2 = 0; w2 < kNmsWarps; ++w2) tot += s_warp_digit[w2][tid];
        s_digit_off[tid] += tot;
      }
      if (valid) { kout[pos] = k; iout[pos] = v; }
      __syncthreads();
    }
    uint32_t* tk = kin; kin = kout; kout = tk;
    int32_t* ti = iin; iin = iout; iout = ti;
  }
  __threadfence_block();
  __syncthreads();

  // ------------------------------- K7b: greedy class-aware suppression -------------------------------
  if (tid == 0) s_kept = 0;
  __syncthreads();
  bool done = false;
  for (int base = 0; base < n && !done; base += kNmsThreads) {
    const int i = base + tid;
    Cand me;
    int my_anchor = -1;
    bool alive = false;
    if (i < n) {
      my_anchor = iin[i];
      alive = eval_anchor(P + static_cast<size_t>(my_anchor) * no, no, thr, me);  // always true
      const int kc = s_kept;
      for (int k = 0; k < kc && alive; ++k)
        if (s_keep[k].cls == me.cls && iou_exceeds(s_keep[k], me, iou_floor)) alive = false;
    }
    // resolve the survivors of this chunk in rank order
    while (true) {
      const uint32_t bal = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) s_warp_first[warp] = bal ? (warp * 32 + __ffs(bal) - 1) : kNmsThreads;
      __syncthreads();
      if (warp == 0) {
        int f = s_warp_first[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
        if (lane == 0) s_first = f;
      }
      __syncthreads();
      const int first = s_first;
      if (first >= kNmsThreads) break;
      const int slot = s_kept;
      if (tid == first) {
        s_keep[slot] = me;
        s_keep_anchor[slot] = my_anchor;
        alive = false;
      }
      __syncthreads();
      if (tid == 0) s_kept = slot + 1;
      if (slot + 1 >= max_det) { done = true; break; }
      if (alive && tid > first && s_keep[slot].cls == me.cls && iou_exceeds(s_keep[slot], me, iou_floor)) alive = false;
      __syncthreads();
    }
    __syncthreads();
  }
  __syncthreads();

  // ------------------------------- K8: rescale + write -------------------------------
  const int kept = min(s_kept, max_det);
  const NmsImageParams ip = params ? params[img] : NmsImageParams{0.f, 0.f, 0.f, 0.f, 0.f};
  float* D = det_out + static_cast<size_t>(img) * max_det * 6;
  for (int k = tid; k < kept; k += kNmsThreads) {
    const Cand c = s_keep[k];
    float x1 = c.x1, y1 = c.y1, x2 = c.x2, y2 = c.y2;
    if (ip.gain != 0.0f) {
      // yolov5 scale_coords + clip_coords, then .round() (pytorch_detector.py:1379)
      x1 = __fdiv_rn(__fsub_rn(x1, ip.pad_x), ip.gain);
      x2 = __fdiv_rn(__fsub_rn(x2, ip.pad_x), ip.gain);
      y1 = __fdiv_rn(__fsub_rn(y1, ip.pad_y), ip.gain);
      y2 = __fdiv_rn(__fsub_rn(y2, ip.pad_y), ip.gain);
      x1 = rintf(fminf(fmaxf(x1, 0.0f), ip.max_x));
      x2 = rintf(fminf(fmaxf(x2, 0.0f), ip.max_x));
      y1 = rintf(fminf(fmaxf(y1, 0.0f), ip.max_y));
      y2 = rintf(fminf(fmaxf(y2, 0.0f), ip.max_y));
    }
    D[k * 6 + 0] = x1; D[k * 6 + 1] = y1; D[k * 6 + 2] = x2; D[k * 6 + 3] = y2;
    D[k * 6 + 4] = c.conf;
    D[k * 6 + 5] = static_cast<float>(c.cls);
    if (kept_anchor) kept_anchor[static_cast<size_t>(img) * max_det + k] = s_keep_anchor[k];
  }
  if (tid == 0) det_count[img] = kept;
}

void nms_launch(const float* pred, int n_img, int anchors, int no, float conf_thres, float iou_thres_floor,
                int max_det, const NmsImageParams* d_params, const NmsScratch& scratch, float* det_out,
                int32_t* det_count, int32_t* kept_anchor, cudaStream_t st) {
  if (n_img <= 0) return;
  MDB_REQUIRE(max_det >= 1 && max_det <= kMaxDetCap, "max_det out of range");
  MDB_REQUIRE(scratch.cap >= anchors, "NMS scratch too small");
  nms_image_kernel<<<n_img, kNmsThreads, 0, st>>>(pred, anchors, no, conf_thres, iou_thres_floor, max_det, d_params,
                                                  scratch, det_out, det_count, kept_anchor);
  MDB_CUDA(cudaGetLastError());
}

}  // namespace mdb
