// K6-K8: confidence threshold + best-class selection + per-class greedy NMS + top-k + box rescale.
//
// Replaces the reference's `nms()` (megadetector/detection/pytorch_detector.py:502-610, a loop of ATen
// kernels, torchvision.ops.nms and >= 3 host syncs per image) and `scale_coords(...).round()`
// (pytorch_detector.py:1379).  Nothing here talks to the host.
//
//   K6  candidate pass, GRID-WIDE (HBM rate): one thread per anchor of every image: obj > thr, conf = p_c*obj,
//       first-max class, conf > thr  ->  64-bit key  (~bits(conf) << 32 | anchor)  appended to the image's candidate
//       list (warp-aggregated atomics; the order of the list is irrelevant because the anchor is part of the key).
//       In the engine this pass is the tail of the Detect decode kernel (aux.cu), i.e. the prediction tensor is
//       thresholded by the thread that produced the row; `cand_from_pred_kernel` is the same pass over a given tensor.
//   K7a one CTA per image: LSD radix sort of the keys, ascending == confidence descending, equal confidence: lower
//       anchor first (torchvision's stable order).  Lists of up to 8192 candidates are sorted entirely in shared memory;
//       longer ones are cut along a cumulative confidence histogram into rounds of <= 16384 keys, highest confidences
//       first (sorted between shared memory and a global partner buffer), and stop as soon as max_det boxes are kept; only
//       a tie group longer than a round falls back to a global sort.  Passes whose digit is constant are skipped.
//   K7b greedy class-aware suppression over the sorted list in windows of 512 candidates: (1) every candidate of the
//       window is tested against the boxes kept so far by two threads (half of the kept list each), (2) the survivors
//       are compacted, (3) their pairwise suppression bit-matrix is computed by all threads, (4) ONE warp resolves the
//       window serially on the bit-matrix (a register OR per survivor, no block barrier).  Because a box can only be
//       suppressed by a higher-ranked box of its own class, walking the merged list and stopping after max_det
//       survivors yields exactly the reference's "per-class NMS, concat, sort by conf, keep 300".
//   K8  scale_coords + clip + round-half-even on the survivors, written with their anchor indices.
//
// Round 1 did all of this in one 1024-thread CTA per image (ordered compaction: 60 barrier rounds; one barrier round
// per survivor): 0.41 ms per launch regardless of the image count, 1.6 % of the HBM rate on the prediction tensor
// (profiles/r01_ncu_nms_details.txt).
//
// Bit-exactness: every float op that torch/torchvision perform is issued with an explicit round-to-nearest
// intrinsic, so nvcc cannot contract a*b+c into an FMA; thresholds are compared in the reference's types
// (conf: float32(thr); IoU: float promoted to double vs the double threshold == float > floor_f32(thr)).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "nms_cand.cuh"

namespace mdb {

// phase cycle counters of the per-image kernel exist only in -DMDB_NMS_TIMERS builds (MDB_LIB_VARIANT experiment builds)
#ifdef MDB_NMS_TIMERS
#define NMS_T(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); s_tm[i] += _t - s_t0; s_t0 = _t; } } while (0)
#else
#define NMS_T(i) do {} while (0)
#endif
constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kWin = 512;                  // candidates per suppression window
constexpr int kWinWords = kWin / 32;       // words per row of the window's suppression bit-matrix
constexpr int kClsLists = 4;               // per-class lists of kept boxes (models with more classes scan the whole kept list)
constexpr int kSortSmemKeys = 8192;        // keys sorted per round in shared memory
constexpr int kBins = 8192;                // confidence bins (2^15 float bit patterns each) used to cut long lists into rounds
constexpr int kBinShift = 15;
// dynamic shared memory: two key buffers (sort ping-pong; after the sort the spare one holds the window state) and the
// cumulative confidence histogram of long lists
constexpr size_t kNmsDynSmem = 2 * static_cast<size_t>(kSortSmemKeys) * sizeof(unsigned long long) + kBins * sizeof(uint32_t);

float float_floor_of_double(double t) {
  float f = static_cast<float>(t);
  if (static_cast<double>(f) > t) f = std::nextafterf(f, -INFINITY);
  return f;
}

size_t nms_scratch_bytes(int n_img, int anchors) {
  return static_cast<size_t>(n_img) * anchors * 16 + static_cast<size_t>(n_img) * sizeof(int32_t) + 256 +
         static_cast<size_t>(n_img) * 8 * sizeof(long long);
}

void nms_bind_scratch(NmsScratch* s, void* mem, int n_img, int anchors) {
  char* p = static_cast<char*>(mem);
  const size_t one = static_cast<size_t>(n_img) * anchors * 8;
  s->key[0] = reinterpret_cast<unsigned long long*>(p);
  s->key[1] = reinterpret_cast<unsigned long long*>(p + one);
  s->count = reinterpret_cast<int32_t*>(p + 2 * one);
  s->dbg = reinterpret_cast<long long*>(p + 2 * one + ((static_cast<size_t>(n_img) * sizeof(int32_t) + 255) / 256) * 256);
  s->cap = anchors;
}

struct Cand {
  float x1, y1, x2, y2, area, conf;
  int cls;
};
static_assert(sizeof(Cand) == 28, "Cand is copied as 7 words");

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
  // disjoint boxes: inter == 0, ovr is 0 (or NaN for two empty boxes) and never exceeds a non-negative threshold -- the
  // common case by far, decided without the division
  if ((xx2 <= xx1 || yy2 <= yy1) && thr_floor >= 0.0f) return false;
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
  return ovr > thr_floor;
}

// ------------------------------- K6 over a given prediction tensor -------------------------------
__global__ void cand_from_pred_kernel(const float* __restrict__ pred, int anchors, int no, float thr,
                                      unsigned long long* __restrict__ keys, int32_t* __restrict__ counts, int cap,
                                      long long total_rows) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool in = r < total_rows;
  const int n = in ? static_cast<int>(r / anchors) : -1;
  const int row = in ? static_cast<int>(r - static_cast<long long>(n) * anchors) : 0;
  float conf = 0.f;
  bool ok = false;
  if (in) {
    const float* p = pred + r * no;
    const float obj = __ldg(p + 4);
    if (obj > thr) {
      float best = -INFINITY;
      for (int k = 0; k < no - 5; ++k) {
        const float cc = __fmul_rn(__ldg(p + 5 + k), obj);
        if (cc > best) best = cc;
      }
      ok = best > thr;
      conf = best;
    }
  }
  cand_append(ok, n, row, conf, keys, counts, cap);
}

// ------------------------------- radix sort of 64-bit keys (one CTA) -------------------------------
// Stable LSD passes over 8 digits of 8 bits; a pass whose digit is the same for all keys moves nothing and is skipped.
// kin/kout may live in shared or in global memory.  Returns the buffer that holds the sorted keys.
__device__ unsigned long long* radix_sort_keys(unsigned long long* kin, unsigned long long* kout, int n,
                                               uint32_t* s_digit_off, uint32_t (*s_warp_digit)[256], int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int pass = 0; pass < 8 && n > 1; ++pass) {
    const int shift = pass * 8;
    if (tid < 256) s_digit_off[tid] = 0;
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kNmsThreads) {
      const int i = base + tid;
      const bool valid = i < n;
      const uint32_t d = valid ? static_cast<uint32_t>(kin[i] >> shift) & 255u : 256u;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (valid && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_digit_off[d], static_cast<uint32_t>(__popc(peers)));
    }
    __syncthreads();
    if (tid < 256 && s_digit_off[tid] == static_cast<uint32_t>(n)) *s_flag = 1;  // constant digit
    __syncthreads();
    const int skip = *s_flag;
    __syncthreads();  // everyone has read the flag before the next pass resets it
    if (skip) continue;
    // exclusive scan over the 256 digit counts (warp 0, 8 per lane)
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
      unsigned long long k = 0;
      uint32_t d = 256u;
      if (valid) { k = kin[i]; d = static_cast<uint32_t>(k >> shift) & 255u; }
      for (int j = lane; j < 256; j += 32) s_warp_digit[warp][j] = 0;
      __syncwarp();
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      if (valid && rank == 0) s_warp_digit[warp][d] = static_cast<uint32_t>(__popc(peers));
      __syncthreads();
      // per digit: exclusive scan of the warps' counts on top of the running digit offset (one thread per digit)
      if (tid < 256) {
        uint32_t run = s_digit_off[tid];
        for (int w2 = 0; w2 < kNmsWarps; ++w2) {
          const uint32_t c = s_warp_digit[w2][tid];
          s_warp_digit[w2][tid] = run;
          run += c;
        }
        s_digit_off[tid] = run;
      }
      __syncthreads();
      if (valid) kout[s_warp_digit[warp][d] + rank] = k;
      __syncthreads();
    }
    unsigned long long* t = kin; kin = kout; kout = t;
  }
  __threadfence_block();
  __syncthreads();
  return kin;
}

// ------------------------------- K7 + K8 -------------------------------
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_image_kernel(const float* __restrict__ pred, int anchors, int no, float thr, float iou_floor, int max_det,
                 const NmsImageParams* __restrict__ params, NmsScratch sc, float* __restrict__ det_out,
                 int32_t* __restrict__ det_count, int32_t* __restrict__ kept_anchor) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ uint32_t s_digit_off[256];
  __shared__ uint32_t s_warp_digit[kNmsWarps][256];
  __shared__ uint32_t s_warp_cnt[kNmsWarps];
  __shared__ int s_flag, s_kept, s_A;
  __shared__ Cand s_keep[kMaxDetCap];
  __shared__ int s_keep_anchor[kMaxDetCap];
  // kept boxes by class: a candidate can only be suppressed by a kept box of its own class, so it is tested against that
  // class's list only (a third of the pair tests for MegaDetector's three classes)
  __shared__ uint16_t s_cls_list[kClsLists][kMaxDetCap];
  __shared__ int s_cls_count[kClsLists];

#ifdef MDB_NMS_TIMERS
  __shared__ long long s_tm[8];
  __shared__ long long s_t0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) s_tm[i] = 0; s_t0 = clock64(); }
#endif
  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = pred + static_cast<size_t>(img) * anchors * no;
  const int n = min(sc.count[img], sc.cap);
  unsigned long long* g0 = sc.key[0] + static_cast<size_t>(img) * sc.cap;
  unsigned long long* g1 = sc.key[1] + static_cast<size_t>(img) * sc.cap;
  unsigned long long* sm0 = reinterpret_cast<unsigned long long*>(s_dyn);
  unsigned long long* sm1 = sm0 + kSortSmemKeys;

  uint32_t* s_cum = reinterpret_cast<uint32_t*>(s_dyn + 2 * static_cast<size_t>(kSortSmemKeys) * sizeof(unsigned long long));
  if (tid == 0) s_kept = 0;
  if (tid < kClsLists) s_cls_count[tid] = 0;
  __syncthreads();
  const bool by_class = (no - 5) <= kClsLists;

  // K7b on `m` sorted keys (window state lives in `spare`, >= 51 KB of shared memory that does not hold the keys);
  // continues from the boxes kept so far, returns when the keys are exhausted or max_det boxes are kept
  auto suppress_sorted = [&](const unsigned long long* sorted, int m, unsigned char* spare) {
    Cand* s_win = reinterpret_cast<Cand*>(spare);                                          // 512 x 28 B = 14336
    int* s_win_anchor = reinterpret_cast<int*>(spare + 14336);                              // 2048
    uint8_t* s_dead1 = reinterpret_cast<uint8_t*>(spare + 14336 + 2048);                     // 512
    uint16_t* s_alist = reinterpret_cast<uint16_t*>(spare + 14336 + 2048 + 512);            // 1024
    uint32_t* s_sup = reinterpret_cast<uint32_t*>(spare + 14336 + 2048 + 512 + 1024);       // 512 x 16 x 4 = 32768
    const int t = tid & (kWin - 1), half = tid >> 9;
    for (int base = 0; base < m; base += kWin) {
      const int i = base + t;
      const bool valid = i < m;
      if (half == 0) {
        Cand c;
        c.cls = -1;
        int anc = -1;
        if (valid) {
          anc = static_cast<int>(static_cast<uint32_t>(sorted[i]));
          eval_anchor(P + static_cast<size_t>(anc) * no, no, thr, c);  // always a candidate
        }
        s_win[t] = c;
        s_win_anchor[t] = anc;
      }
      __syncthreads();
      NMS_T(2);
      const Cand me = s_win[t];
      const int kc = s_kept;
      // (1) against the boxes kept so far: each half of the block scans half of the list
      bool dead = !valid;
      if (valid && by_class) {
        const int cc = s_cls_count[me.cls];
        const uint16_t* lst = s_cls_list[me.cls];
        const int k0 = half ? (cc >> 1) : 0, k1 = half ? cc : (cc >> 1);
        for (int k = k0; k < k1 && !dead; ++k)
          if (iou_exceeds(s_keep[lst[k]], me, iou_floor)) dead = true;
      } else if (valid) {
        const int k0 = half ? (kc >> 1) : 0, k1 = half ? kc : (kc >> 1);
        for (int k = k0; k < k1 && !dead; ++k)
          if (s_keep[k].cls == me.cls && iou_exceeds(s_keep[k], me, iou_floor)) dead = true;
      }
      if (half == 1) s_dead1[t] = dead ? 1 : 0;
      __syncthreads();
      NMS_T(3);
      // (2) compact the survivors of the window, in rank order
      bool alive = false;
      if (half == 0) {
        alive = !dead && !s_dead1[t];
        const uint32_t bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
      }
      __syncthreads();
      if (half == 0) {
        uint32_t before = 0, total = 0;
        for (int w2 = 0; w2 < kWin / 32; ++w2) {
          const uint32_t cnt = s_warp_cnt[w2];
          if (w2 < warp) before += cnt;
          total += cnt;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, alive);
        if (alive) s_alist[before + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint16_t>(t);
        if (tid == 0) s_A = static_cast<int>(total);
      }
      __syncthreads();
      const int A = s_A;
      if (A > 0) {
        // (3) sup[i][w]: bit b set <=> survivor j = 32*w + b ranks after i, has i's class and overlaps i beyond the threshold
        const int wpr = (A + 31) >> 5;
        for (int idx = tid; idx < A * wpr; idx += kNmsThreads) {
          const int ri = idx / wpr, w = idx - ri * wpr;
          uint32_t bits = 0;
          if (w * 32 + 31 > ri) {
            const Cand ci = s_win[s_alist[ri]];
            const int j0 = w * 32;
#pragma unroll 4
            for (int b = 0; b < 32; ++b) {
              const int j = j0 + b;
              if (j > ri && j < A) {
                const Cand cj = s_win[s_alist[j]];
                if (cj.cls == ci.cls && iou_exceeds(ci, cj, iou_floor)) bits |= 1u << b;
              }
            }
          }
          s_sup[ri * kWinWords + w] = bits;
        }
        __syncthreads();
        NMS_T(4);
        // (4) one warp resolves the window: lane l owns word l of the "removed" bit-vector
        if (warp == 0) {
          uint32_t removed = 0;
          int kept = kc;
          uint32_t row = lane < wpr ? s_sup[lane] : 0u;
          for (int ri = 0; ri < A; ++ri) {
            const uint32_t next = (ri + 1 < A && lane < wpr) ? s_sup[(ri + 1) * kWinWords + lane] : 0u;  // independent of the decision
            const uint32_t r = __shfl_sync(0xffffffffu, removed, ri >> 5);
            if (!((r >> (ri & 31)) & 1u)) {
              const int src = s_alist[ri];
              if (lane < 7) reinterpret_cast<uint32_t*>(&s_keep[kept])[lane] = reinterpret_cast<const uint32_t*>(&s_win[src])[lane];
              if (lane == 7) s_keep_anchor[kept] = s_win_anchor[src];
              if (lane == 8 && by_class) {
                const int c = s_win[src].cls;
                s_cls_list[c][s_cls_count[c]++] = static_cast<uint16_t>(kept);
              }
              ++kept;
              if (kept >= max_det) break;
              removed |= row;
            }
            row = next;
          }
          if (lane == 0) s_kept = kept;
        }
      }
      __syncthreads();
      NMS_T(5);
      if (s_kept >= max_det) break;
    }
    __syncthreads();
  };

  if (n <= kSortSmemKeys) {
    // ------------------------------- short list: one round, everything in shared memory -------------------------------
    for (int i = tid; i < n; i += kNmsThreads) sm0[i] = g0[i];
    __syncthreads();
    unsigned long long* res = radix_sort_keys(sm0, sm1, n, s_digit_off, s_warp_digit, &s_flag);
    NMS_T(1);
    suppress_sorted(res, n, reinterpret_cast<unsigned char*>(res == sm0 ? sm1 : sm0));
  } else {
    // ------------------------------- long list: rounds of <= 8192 keys, highest confidences first -------------------------------
    // cumulative histogram over confidence bins (ascending key == descending confidence)
    constexpr uint32_t kBase = 0xC07FFFFFu;  // ~bits(1.0f): the smallest key high word
    auto bin_of = [&](unsigned long long k) -> uint32_t {
      const uint32_t hi = static_cast<uint32_t>(k >> 32);
      const uint32_t b = (hi > kBase ? hi - kBase : 0u) >> kBinShift;
      return b < static_cast<uint32_t>(kBins - 1) ? b : static_cast<uint32_t>(kBins - 1);
    };
    for (int i = tid; i < kBins; i += kNmsThreads) s_cum[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kNmsThreads) atomicAdd(&s_cum[bin_of(g0[i])], 1u);
    __syncthreads();
    {  // inclusive scan of 8192 counters: 8 per thread, warp scan, block scan through s_warp_cnt
      uint32_t loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = s_cum[tid * 8 + j]; sum += loc[j]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp_cnt[warp] = incl;
      __syncthreads();
      uint32_t before = 0;
      for (int w2 = 0; w2 < warp; ++w2) before += s_warp_cnt[w2];
      uint32_t run = before + incl - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) { run += loc[j]; s_cum[tid * 8 + j] = run; }
      __syncthreads();
    }
    // Rounds of up to 16384 keys: the round's keys are gathered into the whole 128 KB key area, radix-sorted with the global
    // buffer g1 as ping-pong partner, left in g1, and the suppression windows then use the key area for their own state.
    constexpr int kRound = 2 * kSortSmemKeys;
    int lo = -1;  // bins <= lo have been processed
    bool fallback = false;
    while (lo < kBins - 1 && s_kept < max_det) {
      const uint32_t done = lo >= 0 ? s_cum[lo] : 0u;
      if (done >= static_cast<uint32_t>(n)) break;
      // largest hi with cum[hi] - done <= kRound (binary search; every thread computes the same value)
      int a = lo + 1, b = kBins - 1;
      if (s_cum[a] - done > static_cast<uint32_t>(kRound)) { fallback = true; break; }  // one bin alone is too long
      while (a < b) {
        const int mid = (a + b + 1) >> 1;
        if (s_cum[mid] - done <= static_cast<uint32_t>(kRound)) a = mid; else b = mid - 1;
      }
      const int hi = a;
      const int m = static_cast<int>(s_cum[hi] - done);
      if (m > 0) {
        if (tid == 0) s_A = 0;
        __syncthreads();
        for (int base = 0; base < n; base += kNmsThreads) {
          const int i = base + tid;
          unsigned long long k = 0;
          bool take = false;
          if (i < n) {
            k = g0[i];
            const int bb = static_cast<int>(bin_of(k));
            take = bb > lo && bb <= hi;
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, take);
          int wbase = 0;
          if (lane == 0 && bal) wbase = atomicAdd(&s_A, __popc(bal));
          wbase = __shfl_sync(0xffffffffu, wbase, 0);
          if (take) sm0[wbase + __popc(bal & ((1u << lane) - 1u))] = k;
        }
        __syncthreads();
        NMS_T(0);
        unsigned long long* res = radix_sort_keys(sm0, g1, m, s_digit_off, s_warp_digit, &s_flag);
        if (res == sm0) {
          for (int i = tid; i < m; i += kNmsThreads) g1[i] = sm0[i];
          __threadfence_block();
          __syncthreads();
        }
        NMS_T(1);
        suppress_sorted(g1, m, s_dyn);
      }
      lo = hi;
      __syncthreads();
    }
    if (fallback) {
      // a single confidence bin holds more than 16384 candidates (mass ties): sort the whole list in global memory and
      // continue after the keys already processed (the rounds above consumed exactly the first cum[lo] sorted keys)
      const uint32_t done = lo >= 0 ? s_cum[lo] : 0u;
      __syncthreads();
      const unsigned long long* sorted = radix_sort_keys(g0, g1, n, s_digit_off, s_warp_digit, &s_flag);
      NMS_T(1);
      suppress_sorted(sorted + done, n - static_cast<int>(done), s_dyn);
    }
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
#ifdef MDB_NMS_TIMERS
  __syncthreads();
  NMS_T(6);
  if (tid == 0) {
    for (int i = 0; i < 7; ++i) sc.dbg[img * 8 + i] = s_tm[i];
    sc.dbg[img * 8 + 7] = n;
  }
#endif
}

void nms_init_device() {
  MDB_CUDA(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kNmsDynSmem)));
}

void nms_candidates_launch(const float* pred, int n_img, int anchors, int no, float conf_thres, const NmsScratch& scratch,
                           cudaStream_t st) {
  if (n_img <= 0) return;
  MDB_REQUIRE(scratch.cap >= anchors, "NMS scratch too small");
  MDB_CUDA(cudaMemsetAsync(scratch.count, 0, sizeof(int32_t) * n_img, st));
  const long long total = static_cast<long long>(n_img) * anchors;
  const int threads = 256;
  cand_from_pred_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, st>>>(
      pred, anchors, no, conf_thres, scratch.key[0], scratch.count, scratch.cap, total);
  MDB_CUDA(cudaGetLastError());
}

void nms_launch(const float* pred, int n_img, int anchors, int no, float conf_thres, float iou_thres_floor,
                int max_det, const NmsImageParams* d_params, const NmsScratch& scratch, float* det_out,
                int32_t* det_count, int32_t* kept_anchor, cudaStream_t st) {
  if (n_img <= 0) return;
  MDB_REQUIRE(max_det >= 1 && max_det <= kMaxDetCap, "max_det out of range");
  MDB_REQUIRE(scratch.cap >= anchors, "NMS scratch too small");
  nms_image_kernel<<<n_img, kNmsThreads, kNmsDynSmem, st>>>(pred, anchors, no, conf_thres, iou_thres_floor, max_det,
                                                            d_params, scratch, det_out, det_count, kept_anchor);
  MDB_CUDA(cudaGetLastError());
#ifdef MDB_NMS_TIMERS
  {
    static int prints = 0;
    if (prints < 6 && getenv("MDB_NMS_DEBUG")) {
      ++prints;
      std::vector<long long> h(static_cast<size_t>(n_img) * 8);
      MDB_CUDA(cudaStreamSynchronize(st));
      MDB_CUDA(cudaMemcpy(h.data(), scratch.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      int worst = 0;
      long long wt = -1;
      for (int i = 0; i < n_img; ++i) {
        long long tot = 0;
        for (int k = 0; k < 7; ++k) tot += h[i * 8 + k];
        if (tot > wt) { wt = tot; worst = i; }
      }
      fprintf(stderr, "[nms dbg] %d images; slowest image %d: n=%lld cycles: select/hist %lld sort %lld load %lld kept-test %lld compact+matrix %lld resolve %lld tail %lld (total %lld)\n",
              n_img, worst, h[worst * 8 + 7], h[worst * 8 + 0], h[worst * 8 + 1], h[worst * 8 + 2], h[worst * 8 + 3], h[worst * 8 + 4],
              h[worst * 8 + 5], h[worst * 8 + 6], wt);
    }
  }
#endif
}

}  // namespace mdb
