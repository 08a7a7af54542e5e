// libmdb200 engine: weights, per-shape execution plans (static graph + liveness-packed activation arena),
// the batch pipeline and the C ABI of include/mdb200.h.
//
// Host-side mirror of what `PTDetector` does around the model
// (megadetector/detection/pytorch_detector.py:745-1424): letterbox geometry, grouping by processed shape
// (:1226-1233), one forward per group (:1257-1313), nms (:1342), scale_coords + round (:1379).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mdb200.h"
#include "common.h"

using namespace mdb;

namespace {

thread_local std::string g_create_error;

struct ConvWeights {
  __half* w = nullptr;  // device [cout_pad][k][k][cin_pad]
  float* b = nullptr;   // device [cout_pad]
  int cout = 0, cout_pad = 0, cin = 0, cin_pad = 0, k = 0;
};

// ------------------------------------------------------------------------------------------
// graph description with virtual buffers
// ------------------------------------------------------------------------------------------
struct VBuf {
  int H, W, C;
  bool fp32 = false;
  int first = 1 << 30, last = -1;
  size_t offset = 0, bytes = 0;
};
struct VView {
  int buf = -1, c_off = 0, C = 0;
};
enum StepKind { STEP_CONV, STEP_UPSAMPLE, STEP_SPPF };
struct Step {
  StepKind kind = STEP_CONV;
  std::string name;
  VView in, out, res;
  int cout = 0, k = 1, s = 1, pad = 0, act = 1;
  int kw = 0, pad_x = -1;  // non-square filter (stem only)
  ConvOp* op = nullptr;
};

struct Plan {
  int H = 0, W = 0, nb = 0;
  std::vector<VBuf> bufs;
  std::vector<Step> steps;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  int stem_buf = -1;
  int raw_buf[4] = {-1, -1, -1, -1};
  float* pred = nullptr;
  int anchors = 0;
  DecodeLevel lv[4];
  void* nms_mem = nullptr;
  NmsScratch nms_scratch{};
  double macs_per_image = 0;
  std::map<int, cudaGraphExec_t> graphs;  // forward pass per batch size
  // test-time augmentation (mdb_set_augment): concatenated predictions of the three passes and their NMS scratch
  float* pred_tta = nullptr;
  int anchors_tta = 0;
  void* nms_mem_tta = nullptr;
  NmsScratch nms_scratch_tta{};
  // parity harness (mdb_debug_set_trace): dense copies of the traced steps' input / output / residual views
  struct TraceRec {
    size_t off[3] = {0, 0, 0};      // in, out, res
    int shape[3][4] = {{0}};        // N, H, W, C
    bool present[3] = {false, false, false};
    bool out_fp32 = false;
  };
  std::map<std::string, TraceRec> trace;
  void* trace_mem = nullptr;
  size_t trace_cap = 0;
  ~Plan() {
    if (trace_mem) cudaFree(trace_mem);
    if (pred_tta) cudaFree(pred_tta);
    if (nms_mem_tta) cudaFree(nms_mem_tta);
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    for (auto& s : steps)
      if (s.op) conv_free(s.op);
    if (arena) cudaFree(arena);
    if (pred) cudaFree(pred);
    if (nms_mem) cudaFree(nms_mem);
  }
  View view(const VView& v) const {
    const VBuf& b = bufs[v.buf];
    View r;
    r.base = static_cast<char*>(arena) + b.offset;
    r.N = nb; r.H = b.H; r.W = b.W; r.Ctot = b.C; r.c_off = v.c_off; r.C = v.C; r.fp32 = b.fp32;
    return r;
  }
};

struct TableKey {
  int ssize, dsize, vertical;
  bool operator<(const TableKey& o) const {
    return std::tie(ssize, dsize, vertical) < std::tie(o.ssize, o.dsize, o.vertical);
  }
};
struct DevTable {
  int32_t* ofs = nullptr;   // linear: source offsets; area: CSR `first`
  int16_t* coef = nullptr;  // linear only
  int32_t* si = nullptr;    // area: source index per tap
  float* alpha = nullptr;   // area: weight per tap
};

struct ImgIn {
  const uint8_t* ptr;
  int h, w;
  long long stride;
  int src_h, src_w;  // original size for scale_coords
  int index;         // position in the caller's arrays
  bool on_device = false;  // ptr is device memory (JPEG path): K1 reads it in place, nothing is uploaded
  double rp_gain = 0, rp_pad_x = 0, rp_pad_y = 0;  // dict inputs: explicit scale_coords ratio_pad ("modern"); 0 = from the shapes
};

struct Pending {
  int slot;
  Plan* plan;
  std::vector<ImgIn> imgs;
  int job = -1;
};

struct ChunkJob {
  int H = 0, W = 0;  // letterboxed shape (plan key)
  std::vector<int> ids;  // indices into BatchState::all
  bool failed = false;
};

// state of the batch that is currently streaming through the engine (mdb_batch_begin / mdb_batch_next)
struct BatchState {
  bool active = false;
  float thr = 0.f;
  std::vector<ImgIn> all;
  std::vector<LetterboxGeom> geom;
  std::vector<ChunkJob> jobs;
  size_t next_job = 0;
  std::vector<Pending> inflight;
  std::vector<size_t> failed_ready;
  std::vector<int> immediate;  // caller indices whose status was final at begin()
  int slot = 0;
  float* det_out = nullptr;
  int32_t* det_count = nullptr;
  int32_t* kept_anchor = nullptr;
  int32_t* status = nullptr;
};

}  // namespace

struct mdb_engine {
  mdb_config cfg{};
  int num_sms = 148;
  cudaStream_t st_compute = nullptr, st_copy = nullptr, st_post = nullptr;
  cudaEvent_t ev_net_done[2] = {nullptr, nullptr}, ev_decode_done = nullptr, ev_post_done = nullptr;
  cudaEvent_t ev_copy_done[2] = {nullptr, nullptr}, ev_compute_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};
  std::map<std::string, ConvWeights> weights;
  bool weights_loaded = false;
  // per-shape plans, bounded: least recently used plans are evicted (arena, pred, NMS scratch, graphs are freed)
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans;
  std::vector<std::pair<int, int>> plan_lru;  // most recent last
  int max_plans = 6;
  std::map<TableKey, DevTable> tables;
  // staging (double-buffered across chunks)
  uint8_t* d_src[2] = {nullptr, nullptr};
  size_t d_src_bytes[2] = {0, 0};
  uint8_t* h_stage[2] = {nullptr, nullptr};  // page-locked staging for images handed in as PAGEABLE host memory
  size_t h_stage_bytes[2] = {0, 0};
  int stage_threads = 8;
  PreImage* h_pre[2] = {nullptr, nullptr};  // pinned
  PreImage* d_pre[2] = {nullptr, nullptr};
  NmsImageParams* h_nmsp[2] = {nullptr, nullptr};  // pinned
  NmsImageParams* d_nmsp[2] = {nullptr, nullptr};
  float* d_det[2] = {nullptr, nullptr};  // [max_batch][max_det][6], one per pipeline slot
  int32_t* d_cnt[2] = {nullptr, nullptr};
  int32_t* d_anchor[2] = {nullptr, nullptr};
  float* h_det[2] = {nullptr, nullptr};  // pinned mirrors
  int32_t* h_cnt[2] = {nullptr, nullptr};
  int32_t* h_anchor[2] = {nullptr, nullptr};
  cudaEvent_t ev_tt[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
  // what is staged in each slot (for mdb_bench_resident)
  Plan* slot_plan[2] = {nullptr, nullptr};
  int slot_n[2] = {0, 0};
  float slot_max_scale[2] = {1.f, 1.f};
  int last_slot = 0;
  // last-run bookkeeping for debug fetch
  Plan* last_plan = nullptr;
  int last_n = 0;
  float last_ms[4] = {0, 0, 0, 0};
  int64_t launches = 0;
  int pipeline_split = 4;
  bool edge_taper = false;
  bool use_graphs = true;
  bool augment = false;  // TTA for the following infer calls (generate_detections_one_batch(..., augment=True))
  bool last_augment = false;
  bool trace_all = false;
  std::vector<std::string> trace_names;
  BatchState batch;
  JpegDecoder* jpeg = nullptr;
  int last_jpeg_backend = 0;
  std::string err;
};

namespace {

int fail(mdb_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg; else g_create_error = msg;
  return code;
}

// ------------------------------------------------------------------------------------------
// graph builder (order == oracle/yolov5.py::conv_specs)
// ------------------------------------------------------------------------------------------
struct Builder {
  Plan& P;
  const mdb_arch& A;
  const std::map<std::string, ConvWeights>* weights = nullptr;  // to see which fused C3 entry convs exist
  explicit Builder(Plan& p, const mdb_arch& a) : P(p), A(a) {}
  int newbuf(int H, int W, int C, bool fp32 = false) {
    VBuf b; b.H = H; b.W = W; b.C = C; b.fp32 = fp32;
    P.bufs.push_back(b);
    return static_cast<int>(P.bufs.size()) - 1;
  }
  VView whole(int buf) { return VView{buf, 0, P.bufs[buf].C}; }
  VView slice(int buf, int off, int C) { return VView{buf, off, C}; }
  void touch(const VView& v) {
    if (v.buf < 0) return;
    const int i = static_cast<int>(P.steps.size());
    P.bufs[v.buf].first = std::min(P.bufs[v.buf].first, i);
    P.bufs[v.buf].last = std::max(P.bufs[v.buf].last, i);
  }
  void conv(const std::string& name, VView in, VView out, int cout, int k, int s, int pad, int act = 1, VView res = VView{}) {
    Step st; st.kind = STEP_CONV; st.name = name; st.in = in; st.out = out; st.res = res;
    st.cout = cout; st.k = k; st.s = s; st.pad = pad; st.act = act;
    touch(in); touch(out); touch(res);
    P.steps.push_back(st);
  }
  void upsample(VView in, VView out, int idx) {
    Step st; st.kind = STEP_UPSAMPLE; st.name = "upsample." + std::to_string(idx); st.in = in; st.out = out;
    touch(in); touch(out);
    P.steps.push_back(st);
  }
  void sppf(VView cat) {
    Step st; st.kind = STEP_SPPF; st.name = "sppf_pool.11"; st.in = cat; st.out = cat;
    touch(cat);
    P.steps.push_back(st);
  }
  // C3(c1 -> c2, n bottlenecks): returns nothing, writes `out`
  void c3(int idx, VView in, VView out, int c2, int n, bool shortcut) {
    const int h = c2 / 2;
    const VBuf& ib = P.bufs[in.buf];
    const int H = ib.H, W = ib.W;
    const std::string pre = "model." + std::to_string(idx);
    const int cat = newbuf(H, W, 2 * h);
    VView a;
    if (weights != nullptr && weights->count(pre + ".cv12.conv")) {
      // cv1 and cv2 read the same tensor: ONE 1x1 conv with the two weight sets stacked along C_out writes
      // [cv1(x) | cv2(x)] straight into the concat buffer (bit-identical per channel; the input is read once instead
      // of twice).  The bottleneck chain then starts from the buffer's first half, and its last member overwrites
      // that half with m(cv1(x)) -- element-wise in place for n == 1 (the residual is read by the thread that stores).
      a = slice(cat, 0, h);
      conv(pre + ".cv12.conv", in, whole(cat), 2 * h, 1, 1, 0);
    } else {
      const int cur = (n == 0) ? -1 : newbuf(H, W, h);
      a = (n == 0) ? slice(cat, 0, h) : whole(cur);
      conv(pre + ".cv1.conv", in, a, h, 1, 1, 0);
      conv(pre + ".cv2.conv", in, slice(cat, h, h), h, 1, 1, 0);
    }
    const int tmp = n > 0 ? newbuf(H, W, h) : -1;
    for (int j = 0; j < n; ++j) {
      const std::string m = pre + ".m." + std::to_string(j);
      conv(m + ".cv1.conv", a, whole(tmp), h, 1, 1, 0);
      VView nxt = (j == n - 1) ? slice(cat, 0, h) : whole(newbuf(H, W, h));
      conv(m + ".cv2.conv", whole(tmp), nxt, h, 3, 1, 1, 1, shortcut ? a : VView{});
      a = nxt;
    }
    conv(pre + ".cv3.conv", whole(cat), out, c2, 1, 1, 0);
  }
};

void build_graph(Plan& P, const mdb_arch& A, const std::map<std::string, ConvWeights>* weights) {
  Builder B(P, A);
  B.weights = weights;
  const int* c = A.ch;
  const int* nb = A.n_backbone;
  const int nh = A.n_head;
  const int H = P.H, W = P.W;
  auto dim = [&](int s, int* h, int* w) { *h = H / s; *w = W / s; };
  int h2, w2, h4, w4, h8, w8, h16, w16, h32, w32, h64, w64;
  dim(2, &h2, &w2); dim(4, &h4, &w4); dim(8, &h8, &w8); dim(16, &h16, &w16); dim(32, &h32, &w32); dim(64, &h64, &w64);

  P.stem_buf = B.newbuf(h2, w2, 48);
  // concat buffers that several producers write into
  const int cat22 = B.newbuf(h8, w8, 2 * c[2]);    // [up(x20), x4]
  const int cat18 = B.newbuf(h16, w16, 2 * c[3]);  // [up(x16), x6]
  const int cat14 = B.newbuf(h32, w32, 2 * c[4]);  // [up(x12), x8]
  const int cat25 = B.newbuf(h16, w16, 2 * c[2]);  // [x24, x20]
  const int cat28 = B.newbuf(h32, w32, 2 * c[3]);  // [x27, x16]
  const int cat31 = B.newbuf(h64, w64, 2 * c[4]);  // [x30, x12]
  const int sppf_cat = B.newbuf(h64, w64, 2 * c[5]);  // [s1, p1, p2, p3], each c5/2

  // NB: the stem consumes K1's tap-packed space-to-depth image: 3x1 / s1 / pad (1,0) over 64 channels
  // (48 stored; the 64-channel TMA box zero-fills the rest; see preprocess.cu and pack_weights_stem_s2d)
  const int x0 = B.newbuf(h2, w2, c[0]);
  B.conv("model.0.conv", B.whole(P.stem_buf), B.whole(x0), c[0], 3, 1, 1);
  P.steps.back().kw = 1;
  P.steps.back().pad_x = 0;
  const int x1 = B.newbuf(h4, w4, c[1]);
  B.conv("model.1.conv", B.whole(x0), B.whole(x1), c[1], 3, 2, 1);
  const int x2 = B.newbuf(h4, w4, c[1]);
  B.c3(2, B.whole(x1), B.whole(x2), c[1], nb[0], true);
  const int x3 = B.newbuf(h8, w8, c[2]);
  B.conv("model.3.conv", B.whole(x2), B.whole(x3), c[2], 3, 2, 1);
  B.c3(4, B.whole(x3), B.slice(cat22, c[2], c[2]), c[2], nb[1], true);  // x4
  const int x5 = B.newbuf(h16, w16, c[3]);
  B.conv("model.5.conv", B.slice(cat22, c[2], c[2]), B.whole(x5), c[3], 3, 2, 1);
  B.c3(6, B.whole(x5), B.slice(cat18, c[3], c[3]), c[3], nb[2], true);  // x6
  const int x7 = B.newbuf(h32, w32, c[4]);
  B.conv("model.7.conv", B.slice(cat18, c[3], c[3]), B.whole(x7), c[4], 3, 2, 1);
  B.c3(8, B.whole(x7), B.slice(cat14, c[4], c[4]), c[4], nb[3], true);  // x8
  const int x9 = B.newbuf(h64, w64, c[5]);
  B.conv("model.9.conv", B.slice(cat14, c[4], c[4]), B.whole(x9), c[5], 3, 2, 1);
  const int x10 = B.newbuf(h64, w64, c[5]);
  B.c3(10, B.whole(x9), B.whole(x10), c[5], nb[4], true);
  // SPPF
  const int hs = c[5] / 2;
  B.conv("model.11.cv1.conv", B.whole(x10), B.slice(sppf_cat, 0, hs), hs, 1, 1, 0);
  B.sppf(B.whole(sppf_cat));
  const int x11 = B.newbuf(h64, w64, c[5]);
  B.conv("model.11.cv2.conv", B.whole(sppf_cat), B.whole(x11), c[5], 1, 1, 0);
  // head
  B.conv("model.12.conv", B.whole(x11), B.slice(cat31, c[4], c[4]), c[4], 1, 1, 0);  // x12
  B.upsample(B.slice(cat31, c[4], c[4]), B.slice(cat14, 0, c[4]), 13);
  const int x15 = B.newbuf(h32, w32, c[4]);
  B.c3(15, B.whole(cat14), B.whole(x15), c[4], nh, false);
  B.conv("model.16.conv", B.whole(x15), B.slice(cat28, c[3], c[3]), c[3], 1, 1, 0);  // x16
  B.upsample(B.slice(cat28, c[3], c[3]), B.slice(cat18, 0, c[3]), 17);
  const int x19 = B.newbuf(h16, w16, c[3]);
  B.c3(19, B.whole(cat18), B.whole(x19), c[3], nh, false);
  B.conv("model.20.conv", B.whole(x19), B.slice(cat25, c[2], c[2]), c[2], 1, 1, 0);  // x20
  B.upsample(B.slice(cat25, c[2], c[2]), B.slice(cat22, 0, c[2]), 21);
  const int x23 = B.newbuf(h8, w8, c[2]);
  B.c3(23, B.whole(cat22), B.whole(x23), c[2], nh, false);
  B.conv("model.24.conv", B.whole(x23), B.slice(cat25, 0, c[2]), c[2], 3, 2, 1);  // x24
  const int x26 = B.newbuf(h16, w16, c[3]);
  B.c3(26, B.whole(cat25), B.whole(x26), c[3], nh, false);
  B.conv("model.27.conv", B.whole(x26), B.slice(cat28, 0, c[3]), c[3], 3, 2, 1);  // x27
  const int x29 = B.newbuf(h32, w32, c[4]);
  B.c3(29, B.whole(cat28), B.whole(x29), c[4], nh, false);
  B.conv("model.30.conv", B.whole(x29), B.slice(cat31, 0, c[4]), c[4], 3, 2, 1);  // x30
  const int x32 = B.newbuf(h64, w64, c[5]);
  B.c3(32, B.whole(cat31), B.whole(x32), c[5], nh, false);
  // Detect heads: fp32 raw outputs, decoded by K5
  const int no = A.nc + 5;
  const int feats[4] = {x23, x26, x29, x32};
  const int fh[4] = {h8, h16, h32, h64}, fw[4] = {w8, w16, w32, w64};
  for (int l = 0; l < 4; ++l) {
    const int craw = ((3 * no + 3) / 4) * 4;
    P.raw_buf[l] = B.newbuf(fh[l], fw[l], craw, true);
    B.conv("model.33.m." + std::to_string(l), B.whole(feats[l]), B.slice(P.raw_buf[l], 0, 3 * no), 3 * no, 1, 1, 0, 0);
  }
  // keep the stem input and the raw heads alive to the end (debug fetch / decode)
  const int last = static_cast<int>(P.steps.size());
  P.bufs[P.stem_buf].first = 0;
  P.bufs[P.stem_buf].last = last;
  for (int l = 0; l < 4; ++l) P.bufs[P.raw_buf[l]].last = last;
}

void assign_offsets(Plan& P, bool reuse) {
  const int n = static_cast<int>(P.bufs.size());
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) {
    order[i] = i;
    VBuf& b = P.bufs[i];
    b.bytes = ((static_cast<size_t>(P.nb) * b.H * b.W * b.C * (b.fp32 ? 4 : 2) + 1023) / 1024) * 1024 + 1024;
  }
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    if (P.bufs[a].bytes != P.bufs[b].bytes) return P.bufs[a].bytes > P.bufs[b].bytes;
    return a < b;
  });
  std::vector<int> placed;
  size_t total = 0;
  for (int id : order) {
    VBuf& b = P.bufs[id];
    if (b.last < 0) { b.offset = 0; continue; }  // never used
    size_t off = 0;
    if (reuse) {
      // lowest offset that does not collide with any lifetime-overlapping, already placed buffer
      bool moved = true;
      while (moved) {
        moved = false;
        for (int q : placed) {
          const VBuf& o = P.bufs[q];
          const bool life = !(o.last < b.first || b.last < o.first);
          const bool addr = !(o.offset + o.bytes <= off || off + b.bytes <= o.offset);
          if (life && addr) { off = o.offset + o.bytes; moved = true; }
        }
      }
    } else {
      off = total;
    }
    b.offset = off;
    total = std::max(total, off + b.bytes);
    placed.push_back(id);
  }
  P.arena_bytes = total;
}

// A plan may be evicted only when nothing refers to it: not staged in a pipeline slot of the running batch, not the
// target of the debug fetches.  The device is idle when this runs (callers synchronise first).
bool plan_in_use(mdb_engine* e, const Plan* P) {
  for (const Pending& pd : e->batch.inflight)
    if (pd.plan == P) return true;
  return false;
}

// frees least-recently-used plans until at most `keep` remain (plans in use are skipped); returns the number freed
int evict_plans(mdb_engine* e, size_t keep) {
  int freed = 0;
  size_t i = 0;
  while (e->plans.size() > keep && i < e->plan_lru.size()) {
    const auto key = e->plan_lru[i];
    auto it = e->plans.find(key);
    if (it == e->plans.end()) { e->plan_lru.erase(e->plan_lru.begin() + i); continue; }
    Plan* P = it->second.get();
    if (plan_in_use(e, P)) { ++i; continue; }
    cudaDeviceSynchronize();
    for (int s = 0; s < 2; ++s)
      if (e->slot_plan[s] == P) { e->slot_plan[s] = nullptr; e->slot_n[s] = 0; }
    if (e->last_plan == P) e->last_plan = nullptr;
    e->plans.erase(it);
    e->plan_lru.erase(e->plan_lru.begin() + i);
    ++freed;
  }
  return freed;
}

Plan* build_plan(mdb_engine* e, int H, int W);

// Plans are cached per letterboxed shape.  The cache is bounded (mdb_config-independent: MDB_MAX_PLANS, default 6 -- the
// classic letterbox produces few distinct shapes per camera, but a mixed-aspect data set can produce dozens, each with
// an arena sized for max_batch): the least recently used plan is dropped first, and an allocation failure drops every
// idle plan and retries once before the shape group is reported as an inference failure.
Plan* get_plan(mdb_engine* e, int H, int W) {
  auto key = std::make_pair(H, W);
  auto it = e->plans.find(key);
  if (it != e->plans.end()) {
    auto pos = std::find(e->plan_lru.begin(), e->plan_lru.end(), key);
    if (pos != e->plan_lru.end()) e->plan_lru.erase(pos);
    e->plan_lru.push_back(key);
    return it->second.get();
  }
  MDB_REQUIRE(e->weights_loaded, "mdb_load_weights must be called before inference");
  MDB_REQUIRE(H % 64 == 0 && W % 64 == 0 && H >= 64 && W >= 64, "letterboxed shape must be a multiple of 64");
  if (e->plans.size() >= static_cast<size_t>(e->max_plans)) evict_plans(e, static_cast<size_t>(e->max_plans) - 1);
  Plan* P = nullptr;
  try {
    P = build_plan(e, H, W);
  } catch (const CudaError&) {
    cudaGetLastError();
    if (evict_plans(e, 0) == 0) throw;
    P = build_plan(e, H, W);  // second failure propagates
  }
  e->plan_lru.push_back(key);
  return P;
}

Plan* build_plan(mdb_engine* e, int H, int W) {
  auto key = std::make_pair(H, W);
  std::unique_ptr<Plan> P(new Plan());
  P->H = H; P->W = W; P->nb = e->cfg.max_batch;
  build_graph(*P, e->cfg.arch, &e->weights);
  assign_offsets(*P, true);
  MDB_CUDA(cudaMalloc(&P->arena, P->arena_bytes));
  MDB_CUDA(cudaMemsetAsync(P->arena, 0, P->arena_bytes, e->st_compute));
  const int no = e->cfg.arch.nc + 5;
  // conv ops
  for (auto& s : P->steps) {
    if (s.kind != STEP_CONV) continue;
    auto wi = e->weights.find(s.name);
    MDB_REQUIRE(wi != e->weights.end(), ("missing weights for " + s.name).c_str());
    const ConvWeights& cw = wi->second;
    ConvDesc d;
    d.in = P->view(s.in);
    d.out = P->view(s.out);
    if (s.res.buf >= 0) d.res = P->view(s.res);
    d.wgt = cw.w; d.bias = cw.b; d.cout = cw.cout; d.cout_pad = cw.cout_pad; d.cin_pad = cw.cin_pad;
    d.k = s.k; d.s = s.s; d.pad = s.pad; d.act = s.act; d.kw = s.kw; d.pad_x = s.pad_x;
    MDB_REQUIRE(cw.cin == d.in.C && cw.k == s.k && cw.cout == s.cout, ("weight shape mismatch for " + s.name).c_str());
    s.op = conv_build(d, e->num_sms);
    P->macs_per_image += conv_macs_per_image(s.op);
  }
  // decode levels + pred
  int off = 0;
  const float strides[4] = {8.f, 16.f, 32.f, 64.f};
  for (int l = 0; l < 4; ++l) {
    const VBuf& b = P->bufs[P->raw_buf[l]];
    DecodeLevel& L = P->lv[l];
    L.raw = reinterpret_cast<const float*>(static_cast<char*>(P->arena) + b.offset);
    L.ny = b.H; L.nx = b.W; L.c_stride = b.C; L.stride = strides[l]; L.row_off = off;
    for (int a = 0; a < 3; ++a) { L.anchors[a][0] = e->cfg.arch.anchors_px[l][a][0]; L.anchors[a][1] = e->cfg.arch.anchors_px[l][a][1]; }
    off += 3 * b.H * b.W;
  }
  P->anchors = off;
  MDB_CUDA(cudaMalloc(&P->pred, static_cast<size_t>(P->nb) * P->anchors * no * sizeof(float)));
  MDB_CUDA(cudaMalloc(&P->nms_mem, nms_scratch_bytes(P->nb, P->anchors)));
  nms_bind_scratch(&P->nms_scratch, P->nms_mem, P->nb, P->anchors);
  Plan* raw = P.get();
  e->plans[key] = std::move(P);
  return raw;
}

void launch_step(mdb_engine* e, Plan* P, Step& s, int n, cudaStream_t st);
void run_network_eager(mdb_engine* e, Plan* P, int n, cudaStream_t st) {
  for (auto& s : P->steps) launch_step(e, P, s, n, st);
}

void launch_step(mdb_engine* e, Plan* P, Step& s, int n, cudaStream_t st) {
  switch (s.kind) {
    case STEP_CONV: conv_launch(s.op, n, st); break;
    case STEP_UPSAMPLE: upsample2x_launch(P->view(s.in), P->view(s.out), n, st); break;
    case STEP_SPPF: {
      const VBuf& b = P->bufs[s.in.buf];
      const int hs = b.C / 4;
      View cat = P->view(s.in);
      View s1 = cat, p1 = cat, p2 = cat, p3 = cat;
      s1.c_off = 0; p1.c_off = hs; p2.c_off = 2 * hs; p3.c_off = 3 * hs;
      s1.C = p1.C = p2.C = p3.C = hs;
      sppf_pool_launch(s1, p1, p2, p3, n, st);
      break;
    }
  }
  e->launches++;
}

// Parity harness: eager pass that keeps dense copies of the selected steps' input, output and residual views (taken
// right before / after the step runs, so in-place steps and arena reuse do not matter).
void run_network_traced(mdb_engine* e, Plan* P, int n, cudaStream_t st) {
  auto wanted = [&](const std::string& nm) {
    if (e->trace_all) return true;
    return std::find(e->trace_names.begin(), e->trace_names.end(), nm) != e->trace_names.end();
  };
  auto dense_bytes = [&](const VView& v, int C) {
    const VBuf& b = P->bufs[v.buf];
    return ((static_cast<size_t>(n) * b.H * b.W * C * (b.fp32 ? 4 : 2) + 255) / 256) * 256;
  };
  P->trace.clear();
  size_t need = 0;
  for (auto& s : P->steps) {
    if (!wanted(s.name)) continue;
    Plan::TraceRec r;
    const VView vin = s.kind == STEP_SPPF ? VView{s.in.buf, 0, P->bufs[s.in.buf].C / 4} : s.in;
    const VView views[3] = {vin, s.out, s.res};
    for (int k = 0; k < 3; ++k) {
      if (views[k].buf < 0) continue;
      const VBuf& b = P->bufs[views[k].buf];
      const int C = (k == 1 && s.kind == STEP_CONV) ? s.cout : views[k].C;
      r.present[k] = true;
      r.off[k] = need;
      r.shape[k][0] = n; r.shape[k][1] = b.H; r.shape[k][2] = b.W; r.shape[k][3] = C;
      if (k == 1) r.out_fp32 = b.fp32;
      need += dense_bytes(views[k], C);
    }
    P->trace[s.name] = r;
  }
  if (need > P->trace_cap) {
    if (P->trace_mem) MDB_CUDA(cudaFree(P->trace_mem));
    P->trace_mem = nullptr; P->trace_cap = 0;
    MDB_CUDA(cudaMalloc(&P->trace_mem, need));
    P->trace_cap = need;
  }
  auto copy_view = [&](const VView& v, int C, size_t off) {
    const VBuf& b = P->bufs[v.buf];
    const size_t el = b.fp32 ? 4 : 2;
    const char* src = static_cast<const char*>(P->arena) + b.offset + static_cast<size_t>(v.c_off) * el;
    MDB_CUDA(cudaMemcpy2DAsync(static_cast<char*>(P->trace_mem) + off, C * el, src, static_cast<size_t>(b.C) * el, C * el,
                               static_cast<size_t>(n) * b.H * b.W, cudaMemcpyDeviceToDevice, st));
  };
  for (auto& s : P->steps) {
    auto it = P->trace.find(s.name);
    if (it != P->trace.end()) {
      const Plan::TraceRec& r = it->second;
      const VView vin = s.kind == STEP_SPPF ? VView{s.in.buf, 0, P->bufs[s.in.buf].C / 4} : s.in;
      copy_view(vin, r.shape[0][3], r.off[0]);
      if (r.present[2]) copy_view(s.res, r.shape[2][3], r.off[2]);
    }
    launch_step(e, P, s, n, st);
    if (it != P->trace.end()) copy_view(s.out, it->second.shape[1][3], it->second.off[1]);
  }
}

// The 167 launches of one forward pass are captured once per (plan, batch size) into a CUDA graph and replayed:
// the host then spends ~10 us instead of ~2 ms per chunk on launches, and the kernels run back to back.
void run_network(mdb_engine* e, Plan* P, int n, cudaStream_t st) {
  if (e->trace_all || !e->trace_names.empty()) {
    run_network_traced(e, P, n, st);
    return;
  }
  if (!e->use_graphs) {
    run_network_eager(e, P, n, st);
    return;
  }
  auto it = P->graphs.find(n);
  if (it == P->graphs.end()) {
    // the conv kernels opt into >48 KB of shared memory lazily; run once eagerly so that no attribute call or
    // first-use initialisation lands inside the capture
    run_network_eager(e, P, n, st);
    MDB_CUDA(cudaStreamSynchronize(st));
    const int64_t launches_before = e->launches;
    cudaGraph_t g = nullptr;
    MDB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      run_network_eager(e, P, n, st);
    } catch (...) {
      cudaStreamEndCapture(st, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    MDB_CUDA(cudaStreamEndCapture(st, &g));
    e->launches = launches_before;  // capturing launches nothing
    cudaGraphExec_t ge = nullptr;
    cudaError_t rc = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (rc != cudaSuccess) {
      cudaGetLastError();
      throw CudaError{std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(rc)};
    }
    it = P->graphs.emplace(n, ge).first;
  }
  MDB_CUDA(cudaGraphLaunch(it->second, st));
  e->launches += static_cast<int64_t>(P->steps.size());
}

const DevTable& get_table(mdb_engine* e, int ssize, int dsize, bool vertical) {
  TableKey k{ssize, dsize, vertical ? 1 : 0};
  auto it = e->tables.find(k);
  if (it != e->tables.end()) return it->second;
  std::vector<int32_t> ofs;
  std::vector<int16_t> coef;
  linear_tables(ssize, dsize, vertical, ofs, coef);
  DevTable t;
  MDB_CUDA(cudaMalloc(&t.ofs, ofs.size() * sizeof(int32_t)));
  MDB_CUDA(cudaMalloc(&t.coef, coef.size() * sizeof(int16_t)));
  // tables are tiny and built once per (size pair): a synchronous copy keeps lifetime handling trivial
  MDB_CUDA(cudaMemcpy(t.ofs, ofs.data(), ofs.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  MDB_CUDA(cudaMemcpy(t.coef, coef.data(), coef.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
  return e->tables.emplace(k, t).first->second;
}

// INTER_AREA tap tables (vertical == 2 / 3 in the cache key: area x / area y)
const DevTable& get_area_table(mdb_engine* e, int ssize, int dsize, bool vertical) {
  TableKey k{ssize, dsize, vertical ? 3 : 2};
  auto it = e->tables.find(k);
  if (it != e->tables.end()) return it->second;
  std::vector<int32_t> first, si;
  std::vector<float> alpha;
  area_tables(ssize, dsize, first, si, alpha);
  DevTable t;
  MDB_CUDA(cudaMalloc(&t.ofs, first.size() * sizeof(int32_t)));
  MDB_CUDA(cudaMalloc(&t.si, std::max<size_t>(1, si.size()) * sizeof(int32_t)));
  MDB_CUDA(cudaMalloc(&t.alpha, std::max<size_t>(1, alpha.size()) * sizeof(float)));
  MDB_CUDA(cudaMemcpy(t.ofs, first.data(), first.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  MDB_CUDA(cudaMemcpy(t.si, si.data(), si.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  MDB_CUDA(cudaMemcpy(t.alpha, alpha.data(), alpha.size() * sizeof(float), cudaMemcpyHostToDevice));
  return e->tables.emplace(k, t).first->second;
}

// geometry of one raw image in the engine's compatibility mode
LetterboxGeom geometry_for(const mdb_config& cfg, int h, int w) {
  if (cfg.compat_mode == 1) return letterbox_geometry_modern(h, w, cfg.img_size, cfg.stride, cfg.use_ceil_for_resize != 0);
  return letterbox_geometry(h, w, cfg.img_size, cfg.stride, cfg.letterbox_auto != 0, cfg.letterbox_scaleup != 0);
}

// fills the resize part of a PreImage (tables, mode) for source size (h, w) -> (g.new_h, g.new_w)
void fill_resize(mdb_engine* e, PreImage& pi, int h, int w, const LetterboxGeom& g) {
  pi.mode = 0;
  pi.iscale_x = pi.iscale_y = 0;
  pi.xa_first = pi.xa_si = pi.ya_first = pi.ya_si = nullptr;
  pi.xa_alpha = pi.ya_alpha = nullptr;
  const DevTable& tx = get_table(e, w, g.new_w, false);
  const DevTable& ty = get_table(e, h, g.new_h, true);
  pi.xofs = tx.ofs; pi.xcoef = tx.coef; pi.yofs = ty.ofs; pi.ycoef = ty.coef;
  if (g.interp == 1 && (g.new_w != w || g.new_h != h)) {
    // cv::resize(INTER_AREA): integer scales on both axes take the ResizeAreaFast path
    const double sx = 1.0 / (static_cast<double>(g.new_w) / w), sy = 1.0 / (static_cast<double>(g.new_h) / h);
    const int isx = static_cast<int>(std::lrint(sx)), isy = static_cast<int>(std::lrint(sy));
    if (std::fabs(sx - isx) < 2.220446049250313e-16 && std::fabs(sy - isy) < 2.220446049250313e-16) {
      pi.mode = 2;
      pi.iscale_x = isx;
      pi.iscale_y = isy;
    } else {
      pi.mode = 1;
      const DevTable& ax = get_area_table(e, w, g.new_w, false);
      const DevTable& ay = get_area_table(e, h, g.new_h, true);
      pi.xa_first = ax.ofs; pi.xa_si = ax.si; pi.xa_alpha = ax.alpha;
      pi.ya_first = ay.ofs; pi.ya_si = ay.si; pi.ya_alpha = ay.alpha;
    }
  }
}

void ensure_src(mdb_engine* e, int slot, size_t bytes) {
  if (e->d_src_bytes[slot] >= bytes) return;
  if (e->d_src[slot]) MDB_CUDA(cudaFree(e->d_src[slot]));
  e->d_src[slot] = nullptr;
  const size_t want = bytes + bytes / 4 + 4096;
  MDB_CUDA(cudaMalloc(&e->d_src[slot], want));
  e->d_src_bytes[slot] = want;
}

// One chunk (<= max_batch images of one letterboxed shape).  enqueue_chunk() queues the H2D copies on the copy
// stream and K1..K8 + the result D2H on the compute stream and returns; finish_chunk() waits for it and scatters
// the results.  Two slots => the copies of chunk i+1 overlap the kernels of chunk i.

void stage_compute(mdb_engine* e, Plan* P, int n, int slot, float conf_thres, int stages_mask, bool timing) {
  cudaStream_t st = e->st_compute;
  const int max_det = e->cfg.max_det;
  const int no = e->cfg.arch.nc + 5;
  if (timing) MDB_CUDA(cudaEventRecord(e->ev_tt[slot][0], st));
  if (stages_mask & 1) {
    View stem = P->view(VView{P->stem_buf, 0, 48});
    preprocess_launch(e->d_pre[slot], n, P->H, P->W, static_cast<__half*>(stem.base), e->slot_max_scale[slot], st);
    e->launches++;
  }
  if (timing) MDB_CUDA(cudaEventRecord(e->ev_tt[slot][1], st));
  // test-time augmentation (yolov5 _forward_augment): scales 1, 0.83 (left-right flipped), 0.67 -- the two extra inputs are
  // derived from K1's output and run through the plans of their own (64-aligned) shapes
  static const double kTtaScale[3] = {1.0, 0.83, 0.67};
  static const bool kTtaFlip[3] = {false, true, false};
  Plan* tp[3] = {P, nullptr, nullptr};
  const bool tta = e->augment;
  if (tta) {
    for (int k = 1; k < 3; ++k) {
      const int hk = static_cast<int>(std::ceil(P->H * kTtaScale[k] / 64.0)) * 64, wk = static_cast<int>(std::ceil(P->W * kTtaScale[k] / 64.0)) * 64;
      tp[k] = get_plan(e, hk, wk);
    }
    tp[0] = get_plan(e, P->H, P->W);  // refresh its LRU position
    if (P->pred_tta == nullptr) {
      // _clip_augmented: the first pass loses its coarsest level, the last pass its finest
      P->anchors_tta = (tp[0]->anchors - 3 * tp[0]->lv[3].ny * tp[0]->lv[3].nx) + tp[1]->anchors +
                       (tp[2]->anchors - 3 * tp[2]->lv[0].ny * tp[2]->lv[0].nx);
      MDB_CUDA(cudaMalloc(&P->pred_tta, static_cast<size_t>(P->nb) * P->anchors_tta * no * sizeof(float)));
      MDB_CUDA(cudaMalloc(&P->nms_mem_tta, nms_scratch_bytes(P->nb, P->anchors_tta)));
      nms_bind_scratch(&P->nms_scratch_tta, P->nms_mem_tta, P->nb, P->anchors_tta);
    }
  }
  if (stages_mask & 2) {
    // the Detect heads of this pass overwrite the raw buffers the previous pass's decode reads
    MDB_CUDA(cudaStreamWaitEvent(st, e->ev_decode_done, 0));
    if (tta) {
      const __half* stem0 = static_cast<const __half*>(P->view(VView{P->stem_buf, 0, 48}).base);
      for (int k = 1; k < 3; ++k) {
        __half* stemk = static_cast<__half*>(tp[k]->view(VView{tp[k]->stem_buf, 0, 48}).base);
        tta_scale_launch(stem0, P->H, P->W, stemk, tp[k]->H, tp[k]->W, kTtaScale[k], kTtaFlip[k], n, st);
        e->launches++;
      }
    }
    run_network(e, P, n, st);
    if (tta) {
      run_network(e, tp[1], n, st);
      run_network(e, tp[2], n, st);
    }
  }
  if (timing) MDB_CUDA(cudaEventRecord(e->ev_tt[slot][2], st));
  if (stages_mask & 4) {
    // decode + NMS keep only a few SMs busy (one CTA per image): they run on their own stream so that the next
    // chunk's K1 and convolutions overlap them
    cudaStream_t sp = e->st_post;
    MDB_CUDA(cudaEventRecord(e->ev_net_done[slot], st));
    MDB_CUDA(cudaStreamWaitEvent(sp, e->ev_net_done[slot], 0));
    if (!tta) {
      decode_launch(P->lv, n, no, P->anchors, P->pred, sp, conf_thres, P->nms_scratch.key[0], P->nms_scratch.count,
                    P->nms_scratch.cap);
      MDB_CUDA(cudaEventRecord(e->ev_decode_done, sp));
      nms_launch(P->pred, n, P->anchors, no, conf_thres, float_floor_of_double(static_cast<double>(e->cfg.iou_thres)),
                 max_det, e->d_nmsp[slot], P->nms_scratch, e->d_det[slot], e->d_cnt[slot], e->d_anchor[slot], sp);
      e->launches += 2;  // decode + candidate pass (one kernel), per-image sort + suppression
    } else {
      // three decodes into one concatenated tensor (_descale_pred + _clip_augmented + torch.cat), then one NMS over it
      int base = 0;
      for (int k = 0; k < 3; ++k) {
        DecodeLevel lv[4];
        int cnt = 0, off = 0;
        for (int l = 0; l < 4; ++l) {
          if ((k == 0 && l == 3) || (k == 2 && l == 0)) continue;  // clipped level
          lv[cnt] = tp[k]->lv[l];
          lv[cnt].row_off = off;
          off += 3 * lv[cnt].ny * lv[cnt].nx;
          ++cnt;
        }
        for (; cnt < 4; ++cnt) { lv[cnt] = lv[0]; lv[cnt].row_off = 0x7fffffff; }
        DecodePass dp{base, P->anchors_tta, static_cast<float>(kTtaScale[k]), kTtaFlip[k] ? static_cast<float>(P->W) : 0.0f};
        decode_launch(lv, n, no, off, P->pred_tta, sp, conf_thres, P->nms_scratch_tta.key[0], P->nms_scratch_tta.count,
                      P->nms_scratch_tta.cap, &dp);
        base += off;
      }
      MDB_CUDA(cudaEventRecord(e->ev_decode_done, sp));
      nms_launch(P->pred_tta, n, P->anchors_tta, no, conf_thres, float_floor_of_double(static_cast<double>(e->cfg.iou_thres)),
                 max_det, e->d_nmsp[slot], P->nms_scratch_tta, e->d_det[slot], e->d_cnt[slot], e->d_anchor[slot], sp);
      e->launches += 4;
    }
    e->last_augment = tta;
    MDB_CUDA(cudaMemcpyAsync(e->h_det[slot], e->d_det[slot], sizeof(float) * n * max_det * 6, cudaMemcpyDeviceToHost, sp));
    MDB_CUDA(cudaMemcpyAsync(e->h_cnt[slot], e->d_cnt[slot], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, sp));
    MDB_CUDA(cudaMemcpyAsync(e->h_anchor[slot], e->d_anchor[slot], sizeof(int32_t) * n * max_det, cudaMemcpyDeviceToHost, sp));
    if (timing) MDB_CUDA(cudaEventRecord(e->ev_tt[slot][3], sp));
  } else if (timing) {
    MDB_CUDA(cudaEventRecord(e->ev_tt[slot][3], st));
  }
}

// the event that marks "everything of this slot's chunk is done" (staging buffers reusable, results on the host)
void record_chunk_done(mdb_engine* e, int slot) {
  MDB_CUDA(cudaEventRecord(e->ev_compute_done[slot], e->st_post));
}

void enqueue_chunk(mdb_engine* e, Plan* P, const std::vector<ImgIn>& imgs, const std::vector<LetterboxGeom>& geoms,
                   float conf_thres, int slot) {
  const int n = static_cast<int>(imgs.size());
  size_t total = 0;
  std::vector<size_t> offs(n);
  for (int i = 0; i < n; ++i) {
    offs[i] = total;
    total += ((static_cast<size_t>(imgs[i].h) * imgs[i].w * 3 + 255) / 256) * 256;
  }
  // the previous user of this slot must be done with its staging buffers before we overwrite them
  MDB_CUDA(cudaEventSynchronize(e->ev_compute_done[slot]));
  ensure_src(e, slot, total + 64);
  // Images in PAGEABLE host memory (what PIL / numpy callers hand in): CUDA would stage them through its own small bounce
  // buffer on this thread (measured: 64 images cost 10 ms more than from page-locked memory).  They are copied into the
  // slot's page-locked staging area by a few host threads instead, and uploaded from there by the copy engine.
  std::vector<char> pageable(n, 0);
  size_t stage_total = 0;
  for (int i = 0; i < n; ++i) {
    if (imgs[i].on_device) continue;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, imgs[i].ptr) != cudaSuccess) { cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; }
    if (at.type == cudaMemoryTypeUnregistered) { pageable[i] = 1; stage_total += static_cast<size_t>(imgs[i].h) * imgs[i].w * 3; }
  }
  if (stage_total > 0) {
    if (e->h_stage_bytes[slot] < total) {
      if (e->h_stage[slot]) MDB_CUDA(cudaFreeHost(e->h_stage[slot]));
      e->h_stage[slot] = nullptr; e->h_stage_bytes[slot] = 0;
      const size_t want = total + total / 4 + 4096;
      MDB_CUDA(cudaHostAlloc(&e->h_stage[slot], want, cudaHostAllocDefault));
      e->h_stage_bytes[slot] = want;
    }
    // rows of all pageable images, dealt to the threads in contiguous byte ranges
    const int T = std::max(1, std::min<int>(e->stage_threads, static_cast<int>(stage_total >> 20) + 1));
    auto copy_range = [&](size_t lo, size_t hi) {  // byte range of the concatenation of the pageable images
      size_t pos = 0;
      for (int i = 0; i < n && pos < hi; ++i) {
        if (!pageable[i]) continue;
        const size_t row = static_cast<size_t>(imgs[i].w) * 3, sz = row * imgs[i].h;
        if (pos + sz > lo) {
          const size_t a = std::max(lo, pos) - pos, b = std::min(hi, pos + sz) - pos;
          uint8_t* dst = e->h_stage[slot] + offs[i];
          if (imgs[i].stride == static_cast<long long>(row)) {
            memcpy(dst + a, imgs[i].ptr + a, b - a);
          } else {
            for (size_t r = a / row; r * row < b; ++r) {
              const size_t c0 = std::max(a, r * row) - r * row, c1 = std::min(b, (r + 1) * row) - r * row;
              memcpy(dst + r * row + c0, imgs[i].ptr + static_cast<size_t>(imgs[i].stride) * r + c0, c1 - c0);
            }
          }
        }
        pos += sz;
      }
    };
    if (T == 1) {
      copy_range(0, stage_total);
    } else {
      std::vector<std::thread> th;
      const size_t per = (stage_total + T - 1) / T;
      for (int t = 0; t < T; ++t) th.emplace_back(copy_range, std::min(stage_total, per * t), std::min(stage_total, per * (t + 1)));
      for (auto& x : th) x.join();
    }
  }
  float max_scale = 1.0f;
  for (int i = 0; i < n; ++i) {
    const ImgIn& im = imgs[i];
    uint8_t* d = e->d_src[slot] + offs[i];
    const size_t row = static_cast<size_t>(im.w) * 3;
    if (im.on_device) {
      d = const_cast<uint8_t*>(im.ptr);  // decoded by nvJPEG on the copy stream, ordered before ev_copy_done below
    } else if (pageable[i]) {
      MDB_CUDA(cudaMemcpyAsync(d, e->h_stage[slot] + offs[i], row * im.h, cudaMemcpyHostToDevice, e->st_copy));
    } else if (im.stride == static_cast<long long>(row)) {
      MDB_CUDA(cudaMemcpyAsync(d, im.ptr, row * im.h, cudaMemcpyHostToDevice, e->st_copy));
    } else {
      MDB_CUDA(cudaMemcpy2DAsync(d, row, im.ptr, static_cast<size_t>(im.stride), row, im.h, cudaMemcpyHostToDevice,
                                 e->st_copy));
    }
    const LetterboxGeom& g = geoms[i];
    PreImage& pi = e->h_pre[slot][i];
    pi.src = d; pi.src_stride = static_cast<long long>(row); pi.src_h = im.h; pi.src_w = im.w;
    pi.new_w = g.new_w; pi.new_h = g.new_h; pi.top = g.top; pi.left = g.left;
    fill_resize(e, pi, im.h, im.w, g);
    max_scale = std::max(max_scale, std::max(static_cast<float>(im.w) / g.new_w, static_cast<float>(im.h) / g.new_h));
    // scale_coords(img1_shape=(H,W), coords, img0_shape, ratio_pad) -- yolov5 utils.general; ratio_pad=None in "classic"
    NmsImageParams& np = e->h_nmsp[slot][i];
    double gain = std::min(static_cast<double>(P->H) / im.src_h, static_cast<double>(P->W) / im.src_w);
    double padx = (P->W - im.src_w * gain) / 2, pady = (P->H - im.src_h * gain) / 2;
    if (g.sc_gain > 0) { gain = g.sc_gain; padx = g.sc_pad_x; pady = g.sc_pad_y; }
    np.gain = static_cast<float>(gain); np.pad_x = static_cast<float>(padx); np.pad_y = static_cast<float>(pady);
    np.max_x = static_cast<float>(im.src_w); np.max_y = static_cast<float>(im.src_h);
  }
  MDB_CUDA(cudaMemcpyAsync(e->d_pre[slot], e->h_pre[slot], sizeof(PreImage) * n, cudaMemcpyHostToDevice, e->st_copy));
  MDB_CUDA(cudaMemcpyAsync(e->d_nmsp[slot], e->h_nmsp[slot], sizeof(NmsImageParams) * n, cudaMemcpyHostToDevice, e->st_copy));
  MDB_CUDA(cudaEventRecord(e->ev_copy_done[slot], e->st_copy));
  e->slot_plan[slot] = P;
  e->slot_n[slot] = n;
  e->slot_max_scale[slot] = max_scale;
  e->last_slot = slot;
  MDB_CUDA(cudaStreamWaitEvent(e->st_compute, e->ev_copy_done[slot], 0));
  stage_compute(e, P, n, slot, conf_thres, 7, true);
  record_chunk_done(e, slot);
}

void finish_chunk(mdb_engine* e, const Pending& pd, float* det_out, int32_t* det_count, int32_t* kept_anchor) {
  const int slot = pd.slot;
  const int n = static_cast<int>(pd.imgs.size());
  const int max_det = e->cfg.max_det;
  MDB_CUDA(cudaEventSynchronize(e->ev_compute_done[slot]));
  for (int i = 0; i < n; ++i) {
    const int dst = pd.imgs[i].index;
    det_count[dst] = e->h_cnt[slot][i];
    memcpy(det_out + static_cast<size_t>(dst) * max_det * 6, e->h_det[slot] + static_cast<size_t>(i) * max_det * 6,
           sizeof(float) * max_det * 6);
    if (kept_anchor)
      memcpy(kept_anchor + static_cast<size_t>(dst) * max_det, e->h_anchor[slot] + static_cast<size_t>(i) * max_det,
             sizeof(int32_t) * max_det);
  }
  float a = 0, b = 0, c = 0, d = 0;
  cudaEventElapsedTime(&a, e->ev_tt[slot][0], e->ev_tt[slot][1]);
  cudaEventElapsedTime(&b, e->ev_tt[slot][1], e->ev_tt[slot][2]);
  cudaEventElapsedTime(&c, e->ev_tt[slot][2], e->ev_tt[slot][3]);
  cudaEventElapsedTime(&d, e->ev_tt[slot][0], e->ev_tt[slot][3]);
  e->last_ms[0] += a; e->last_ms[1] += b; e->last_ms[2] += c; e->last_ms[3] += d;
  e->last_plan = pd.plan;
  e->last_n = n;
}

// ---- streaming batch: begin() plans the chunks and queues the first two, next() completes the oldest chunk,
// reports which images are ready and queues one more.  The host can format finished records while the GPU works
// on the following chunks (B200Detector does); mdb_infer_batch is begin + next-until-empty. ----
void fail_job(mdb_engine* e, BatchState& B, size_t job, const std::string& why) {
  e->err = why;
  // a whole shape group fails together (pytorch_detector.py:1236-1249): this job and every later job of its shape
  const int H = B.jobs[job].H, W = B.jobs[job].W;
  for (size_t j = job; j < B.jobs.size(); ++j) {
    if (B.jobs[j].H != H || B.jobs[j].W != W || B.jobs[j].failed) continue;
    B.jobs[j].failed = true;
    for (int id : B.jobs[j].ids) { B.status[B.all[id].index] = MDB_IMG_INFER_FAILED; B.det_count[B.all[id].index] = 0; }
  }
}

void enqueue_next_job(mdb_engine* e, BatchState& B) {
  while (B.next_job < B.jobs.size() && B.inflight.size() < 2) {
    const size_t j = B.next_job++;
    ChunkJob& job = B.jobs[j];
    if (job.failed) { B.failed_ready.push_back(j); continue; }
    try {
      Plan* P = get_plan(e, job.H, job.W);
      Pending pd;
      pd.slot = B.slot;
      pd.plan = P;
      pd.job = static_cast<int>(j);
      std::vector<LetterboxGeom> cg;
      for (int id : job.ids) { pd.imgs.push_back(B.all[id]); cg.push_back(B.geom[id]); }
      enqueue_chunk(e, P, pd.imgs, cg, B.thr, B.slot);
      B.inflight.push_back(pd);
      B.slot ^= 1;
    } catch (const CudaError& ex) {
      fail_job(e, B, j, ex.msg);
      B.failed_ready.push_back(j);
      cudaError_t sticky = cudaGetLastError();
      if (sticky != cudaSuccess && cudaDeviceSynchronize() != cudaSuccess) throw;
    }
  }
}

int batch_begin(mdb_engine* e, std::vector<ImgIn>& all, bool preprocessed, float conf_thres, float* det_out,
                int32_t* det_count, int32_t* kept_anchor, int32_t* status) {
  const mdb_config& cfg = e->cfg;
  BatchState& B = e->batch;
  if (B.active) return fail(e, MDB_ERR_STATE, "a batch is already in progress");
  if (e->tables.size() > 4096) {  // resize-coefficient tables are a few KB each; nothing is in flight here
    cudaDeviceSynchronize();
    for (auto& kv : e->tables) { cudaFree(kv.second.ofs); cudaFree(kv.second.coef); cudaFree(kv.second.si); cudaFree(kv.second.alpha); }
    e->tables.clear();
  }
  B = BatchState();
  B.active = true;
  B.thr = conf_thres;
  B.det_out = det_out; B.det_count = det_count; B.kept_anchor = kept_anchor; B.status = status;
  B.all = all;
  B.geom.resize(all.size());
  memset(e->last_ms, 0, sizeof(e->last_ms));
  // group by letterboxed shape, keeping input order inside a group (pytorch_detector.py:1226-1233)
  std::map<std::pair<int, int>, std::vector<int>> groups;
  for (size_t i = 0; i < all.size(); ++i) {
    const ImgIn& im = all[i];
    if (im.ptr == nullptr || im.h <= 0 || im.w <= 0) {
      status[im.index] = MDB_IMG_BAD_INPUT; det_count[im.index] = 0;
      B.immediate.push_back(im.index);
      continue;
    }
    LetterboxGeom g;
    if (preprocessed) {
      g = LetterboxGeom{}; g.src_h = im.h; g.src_w = im.w; g.new_w = im.w; g.new_h = im.h; g.out_h = im.h; g.out_w = im.w;
      g.sc_gain = im.rp_gain; g.sc_pad_x = im.rp_pad_x; g.sc_pad_y = im.rp_pad_y;
    } else {
      g = geometry_for(cfg, im.h, im.w);
    }
    if (g.out_h % 64 != 0 || g.out_w % 64 != 0 || g.new_w < 1 || g.new_h < 1) {
      status[im.index] = MDB_IMG_INFER_FAILED; det_count[im.index] = 0;
      B.immediate.push_back(im.index);
      continue;
    }
    B.geom[i] = g;
    groups[std::make_pair(g.out_h, g.out_w)].push_back(static_cast<int>(i));
    status[im.index] = MDB_IMG_OK;
  }
  for (auto& kv : groups) {
    const std::vector<int>& ids = kv.second;
    // chunking: max_batch-sized chunks when there are several; a single group that fits one chunk is split
    // into `pipeline_split` parts so that uploads and host-side formatting overlap the kernels
    size_t chunk = static_cast<size_t>(cfg.max_batch);
    if (ids.size() <= chunk && ids.size() > 16 && e->pipeline_split > 1)
      chunk = std::max<size_t>(8, (ids.size() + e->pipeline_split - 1) / e->pipeline_split);
    // Nothing hides the upload of the FIRST chunk or the NMS + D2H + record formatting of the LAST one (the reference's call
    // is synchronous per batch), so with `edge_taper` those two chunks are half-sized: [c/2, c, ..., c, c/2].
    std::vector<size_t> sizes;
    {
      size_t left = ids.size();
      const bool taper = e->edge_taper && chunk >= 16 && ids.size() > 2 * chunk && e->pipeline_split > 1;
      if (taper) { sizes.push_back(chunk / 2); left -= chunk / 2; }
      // (a short head chunk for 6..16-image calls -- GPU busy after a quarter of the upload -- was measured and lost: e2e 1111 /
      // 1134 -> 1010 / 1000 images/s at 8 images; a 2-image pass through 156 launches costs more than the upload it hides)
      while (left > 0) {
        size_t c = std::min(left, chunk);
        if (taper && left > chunk / 2 && left - c < chunk / 2 && left - c > 0) c = left - chunk / 2;  // leave exactly c/2 for the tail
        if (taper && left <= chunk && left > chunk / 2) c = left - chunk / 2;
        sizes.push_back(c);
        left -= c;
      }
    }
    size_t b = 0;
    for (size_t c : sizes) {
      ChunkJob job;
      job.H = kv.first.first; job.W = kv.first.second;
      job.ids.assign(ids.begin() + b, ids.begin() + b + c);
      B.jobs.push_back(job);
      b += c;
    }
  }
  try {
    enqueue_next_job(e, B);
  } catch (const CudaError& ex) {
    B.active = false;
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

// completes the oldest outstanding piece of work; writes the caller-array indices that became final
int batch_next(mdb_engine* e, int32_t* done_idx, int32_t cap, int32_t* n_done) {
  BatchState& B = e->batch;
  *n_done = 0;
  if (!B.active) return MDB_OK;
  auto emit = [&](int idx) { if (*n_done < cap) done_idx[(*n_done)++] = idx; };
  try {
    if (!B.immediate.empty()) {
      for (int idx : B.immediate) emit(idx);
      B.immediate.clear();
      return MDB_OK;
    }
    if (!B.failed_ready.empty()) {
      const size_t j = B.failed_ready.front();
      B.failed_ready.erase(B.failed_ready.begin());
      for (int id : B.jobs[j].ids) emit(B.all[id].index);
      return MDB_OK;
    }
    if (!B.inflight.empty()) {
      Pending pd = B.inflight.front();
      B.inflight.erase(B.inflight.begin());
      try {
        finish_chunk(e, pd, B.det_out, B.det_count, B.kept_anchor);
      } catch (const CudaError& ex) {
        fail_job(e, B, static_cast<size_t>(pd.job), ex.msg);
        cudaError_t sticky = cudaGetLastError();
        if (sticky != cudaSuccess && cudaDeviceSynchronize() != cudaSuccess) throw;
      }
      for (const ImgIn& im : pd.imgs) emit(im.index);
      enqueue_next_job(e, B);
      return MDB_OK;
    }
    enqueue_next_job(e, B);
    if (B.inflight.empty() && B.failed_ready.empty()) B.active = false;
    else return batch_next(e, done_idx, cap, n_done);
  } catch (const CudaError& ex) {
    B.active = false;
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

int infer_common(mdb_engine* e, std::vector<ImgIn>& all, bool preprocessed, float conf_thres, float* det_out,
                 int32_t* det_count, int32_t* kept_anchor, int32_t* status, int n) {
  int rc = batch_begin(e, all, preprocessed, conf_thres, det_out, det_count, kept_anchor, status);
  if (rc != MDB_OK) return rc;
  std::vector<int32_t> done(std::max(1, n));
  while (true) {
    int32_t nd = 0;
    rc = batch_next(e, done.data(), static_cast<int32_t>(done.size()), &nd);
    if (rc != MDB_OK) return rc;
    if (nd == 0) break;
  }
  return MDB_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* mdb_version(void) { return "mdb200 0.1 (sm_100a)"; }

const char* mdb_last_error(const mdb_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int mdb_letterbox_geometry(const mdb_config* cfg, int32_t h, int32_t w, int32_t* out8, double* ratio_pad3) {
  if (!cfg || !out8 || h <= 0 || w <= 0) return MDB_ERR_ARG;
  const LetterboxGeom g = geometry_for(*cfg, h, w);
  out8[0] = g.new_w; out8[1] = g.new_h; out8[2] = g.top; out8[3] = g.bottom; out8[4] = g.left; out8[5] = g.right;
  out8[6] = g.out_h; out8[7] = g.out_w;
  if (ratio_pad3) { ratio_pad3[0] = g.ratio; ratio_pad3[1] = g.dw; ratio_pad3[2] = g.dh; }
  return MDB_OK;
}

void* mdb_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void mdb_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int mdb_create(mdb_engine** out, const mdb_config* cfg) {
  if (!out || !cfg) return fail(nullptr, MDB_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->struct_size != static_cast<int32_t>(sizeof(mdb_config))) return fail(nullptr, MDB_ERR_ARG, "mdb_config size mismatch");
  if (cfg->max_batch < 1 || cfg->max_det < 1 || cfg->max_det > kMaxDetCap) return fail(nullptr, MDB_ERR_ARG, "bad max_batch / max_det");
  // every conv needs C_in % 16 == 0; C3 hidden width is ch/2 for ch[1..5]
  if (cfg->arch.ch[0] < 16 || cfg->arch.ch[0] % 16 != 0) return fail(nullptr, MDB_ERR_ARG, "arch.ch[0] must be a multiple of 16");
  for (int i = 1; i < 6; ++i)
    if (cfg->arch.ch[i] < 32 || cfg->arch.ch[i] % 32 != 0) return fail(nullptr, MDB_ERR_ARG, "arch.ch[1..5] must be multiples of 32");
  if (cfg->arch.nc < 1 || cfg->arch.nc > 251) return fail(nullptr, MDB_ERR_ARG, "arch.nc out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(nullptr, MDB_ERR_CUDA, "no CUDA device: libmdb200 has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MDB_ERR_ARG, "device ordinal out of range");
  mdb_engine* e = new mdb_engine();
  try {
    e->cfg = *cfg;
    MDB_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    MDB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    MDB_REQUIRE(prop.major == 10, "libmdb200 is built for sm_100a (B200) only");
    e->num_sms = prop.multiProcessorCount;
    MDB_CUDA(cudaStreamCreateWithFlags(&e->st_compute, cudaStreamNonBlocking));
    MDB_CUDA(cudaStreamCreateWithFlags(&e->st_copy, cudaStreamNonBlocking));
    MDB_CUDA(cudaStreamCreateWithFlags(&e->st_post, cudaStreamNonBlocking));
    MDB_CUDA(cudaEventCreateWithFlags(&e->ev_decode_done, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&e->ev_post_done, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) MDB_CUDA(cudaEventCreateWithFlags(&e->ev_net_done[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      MDB_CUDA(cudaEventCreateWithFlags(&e->ev_copy_done[i], cudaEventDisableTiming));
      MDB_CUDA(cudaEventCreateWithFlags(&e->ev_compute_done[i], cudaEventDisableTiming));
      MDB_CUDA(cudaEventRecord(e->ev_compute_done[i], e->st_compute));
      MDB_CUDA(cudaHostAlloc(&e->h_pre[i], sizeof(PreImage) * cfg->max_batch, cudaHostAllocDefault));
      MDB_CUDA(cudaMalloc(&e->d_pre[i], sizeof(PreImage) * cfg->max_batch));
      MDB_CUDA(cudaHostAlloc(&e->h_nmsp[i], sizeof(NmsImageParams) * cfg->max_batch, cudaHostAllocDefault));
      MDB_CUDA(cudaMalloc(&e->d_nmsp[i], sizeof(NmsImageParams) * cfg->max_batch));
    }
    for (int i = 0; i < 4; ++i) MDB_CUDA(cudaEventCreate(&e->ev_t[i]));
    const size_t nd = static_cast<size_t>(cfg->max_batch) * cfg->max_det;
    for (int i = 0; i < 2; ++i) {
      MDB_CUDA(cudaMalloc(&e->d_det[i], nd * 6 * sizeof(float)));
      MDB_CUDA(cudaMalloc(&e->d_cnt[i], cfg->max_batch * sizeof(int32_t)));
      MDB_CUDA(cudaMalloc(&e->d_anchor[i], nd * sizeof(int32_t)));
      MDB_CUDA(cudaHostAlloc(&e->h_det[i], nd * 6 * sizeof(float), cudaHostAllocDefault));
      MDB_CUDA(cudaHostAlloc(&e->h_cnt[i], cfg->max_batch * sizeof(int32_t), cudaHostAllocDefault));
      MDB_CUDA(cudaHostAlloc(&e->h_anchor[i], nd * sizeof(int32_t), cudaHostAllocDefault));
      for (int j = 0; j < 4; ++j) MDB_CUDA(cudaEventCreate(&e->ev_tt[i][j]));
    }
    preprocess_init_device();
    conv_init_device();
    nms_init_device();
    if (const char* mp = getenv("MDB_MAX_PLANS")) e->max_plans = std::max(1, atoi(mp));
    if (const char* et = getenv("MDB_EDGE_TAPER")) e->edge_taper = et[0] == '1';
    if (const char* stg = getenv("MDB_STAGE_THREADS")) e->stage_threads = std::max(0, atoi(stg));
    const char* ng = getenv("MDB_NO_GRAPHS");
    e->use_graphs = !(ng && ng[0] == '1');
  } catch (const CudaError& ex) {
    g_create_error = ex.msg;
    mdb_destroy(e);
    return MDB_ERR_CUDA;
  }
  *out = e;
  return MDB_OK;
}

void mdb_destroy(mdb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  e->plans.clear();
  for (auto& kv : e->weights) { cudaFree(kv.second.w); cudaFree(kv.second.b); }
  for (auto& kv : e->tables) { cudaFree(kv.second.ofs); cudaFree(kv.second.coef); cudaFree(kv.second.si); cudaFree(kv.second.alpha); }
  for (int i = 0; i < 2; ++i) {
    if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]);
    if (e->d_src[i]) cudaFree(e->d_src[i]);
    if (e->h_pre[i]) cudaFreeHost(e->h_pre[i]);
    if (e->d_pre[i]) cudaFree(e->d_pre[i]);
    if (e->h_nmsp[i]) cudaFreeHost(e->h_nmsp[i]);
    if (e->d_nmsp[i]) cudaFree(e->d_nmsp[i]);
    if (e->ev_copy_done[i]) cudaEventDestroy(e->ev_copy_done[i]);
    if (e->ev_compute_done[i]) cudaEventDestroy(e->ev_compute_done[i]);
  }
  for (int i = 0; i < 4; ++i) if (e->ev_t[i]) cudaEventDestroy(e->ev_t[i]);
  for (int i = 0; i < 2; ++i) {
    cudaFree(e->d_det[i]); cudaFree(e->d_cnt[i]); cudaFree(e->d_anchor[i]);
    if (e->h_det[i]) cudaFreeHost(e->h_det[i]);
    if (e->h_cnt[i]) cudaFreeHost(e->h_cnt[i]);
    if (e->h_anchor[i]) cudaFreeHost(e->h_anchor[i]);
    for (int j = 0; j < 4; ++j) if (e->ev_tt[i][j]) cudaEventDestroy(e->ev_tt[i][j]);
  }
  if (e->jpeg) jpeg_destroy(e->jpeg);
  if (e->st_compute) cudaStreamDestroy(e->st_compute);
  if (e->st_copy) cudaStreamDestroy(e->st_copy);
  if (e->st_post) cudaStreamDestroy(e->st_post);
  if (e->ev_decode_done) cudaEventDestroy(e->ev_decode_done);
  if (e->ev_post_done) cudaEventDestroy(e->ev_post_done);
  for (int i = 0; i < 2; ++i) if (e->ev_net_done[i]) cudaEventDestroy(e->ev_net_done[i]);
  delete e;
}

int mdb_load_weights(mdb_engine* e, const mdb_tensor_desc* tensors, int32_t n) {
  if (!e || !tensors || n <= 0) return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    std::map<std::string, const mdb_tensor_desc*> byname;
    for (int i = 0; i < n; ++i) byname[tensors[i].name] = &tensors[i];
    for (auto& kv : byname) {
      const std::string& nm = kv.first;
      const std::string suffix = ".weight";
      if (nm.size() <= suffix.size() || nm.compare(nm.size() - suffix.size(), suffix.size(), suffix) != 0) continue;
      const std::string base = nm.substr(0, nm.size() - suffix.size());
      const mdb_tensor_desc* tw = kv.second;
      auto bi = byname.find(base + ".bias");
      MDB_REQUIRE(bi != byname.end(), ("missing bias for " + base).c_str());
      MDB_REQUIRE(tw->ndim == 4 && tw->shape[2] == tw->shape[3], "conv weight must be [co][ci][k][k]");
      const int co = static_cast<int>(tw->shape[0]), ci = static_cast<int>(tw->shape[1]), k = static_cast<int>(tw->shape[2]);
      MDB_REQUIRE(bi->second->ndim == 1 && bi->second->shape[0] == co, "bias shape mismatch");
      ConvWeights cw;
      cw.cout = co;
      cw.cout_pad = ((co + 15) / 16) * 16;
      std::vector<__half> packed;
      if (base == "model.0.conv") {
        MDB_REQUIRE(ci == 3 && k == 6, "stem must be a 6x6 conv over 3 channels");
        cw.cin = 48; cw.cin_pad = 64; cw.k = 3;
        pack_weights_stem_s2d(tw->data, co, cw.cout_pad, cw.cin_pad, packed);
      } else {
        MDB_REQUIRE(ci % 16 == 0, "conv input channels must be a multiple of 16");
        cw.cin = ci; cw.cin_pad = conv_cin_pad(ci); cw.k = k;
        pack_weights_khwc(tw->data, co, ci, k, cw.cout_pad, cw.cin_pad, packed);
      }
      std::vector<float> bias(cw.cout_pad, 0.f);
      memcpy(bias.data(), bi->second->data, sizeof(float) * co);
      auto old = e->weights.find(base);
      if (old != e->weights.end()) { cudaFree(old->second.w); cudaFree(old->second.b); }
      MDB_CUDA(cudaMalloc(&cw.w, packed.size() * sizeof(__half) + 256));
      MDB_CUDA(cudaMalloc(&cw.b, bias.size() * sizeof(float)));
      MDB_CUDA(cudaMemcpy(cw.w, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
      MDB_CUDA(cudaMemcpy(cw.b, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
      e->weights[base] = cw;
    }
    // C3 entry convs: cv1 and cv2 share their input -> one stacked weight set "<block>.cv12.conv" (MDB_FUSE_C3=0 keeps
    // them separate).  Needs cv1's rows unpadded so that cv2's channels follow directly.
    {
      const char* ef = getenv("MDB_FUSE_C3");
      const bool fuse = !(ef && ef[0] == '0');
      std::vector<std::string> bases;
      for (auto& kv : e->weights) {
        const std::string suffix = ".cv1.conv";
        const std::string& nm = kv.first;
        if (nm.size() > suffix.size() && nm.compare(nm.size() - suffix.size(), suffix.size(), suffix) == 0 &&
            nm.find(".m.") == std::string::npos)
          bases.push_back(nm.substr(0, nm.size() - suffix.size()));
      }
      for (const std::string& b : bases) {
        auto old = e->weights.find(b + ".cv12.conv");
        if (old != e->weights.end()) { cudaFree(old->second.w); cudaFree(old->second.b); e->weights.erase(old); }
        auto i2 = e->weights.find(b + ".cv2.conv");
        if (!fuse || i2 == e->weights.end()) continue;
        const ConvWeights& w1 = e->weights[b + ".cv1.conv"];
        const ConvWeights& w2 = i2->second;
        if (w1.k != 1 || w2.k != 1 || w1.cin != w2.cin || w1.cin_pad != w2.cin_pad || w1.cout != w1.cout_pad ||
            w1.cout_pad + w2.cout_pad > 1280)
          continue;
        ConvWeights cw;
        cw.cout = w1.cout + w2.cout; cw.cout_pad = w1.cout_pad + w2.cout_pad; cw.cin = w1.cin; cw.cin_pad = w1.cin_pad; cw.k = 1;
        const size_t row = static_cast<size_t>(cw.cin_pad) * sizeof(__half);
        MDB_CUDA(cudaMalloc(&cw.w, row * cw.cout_pad + 256));
        MDB_CUDA(cudaMalloc(&cw.b, sizeof(float) * cw.cout_pad));
        MDB_CUDA(cudaMemcpy(cw.w, w1.w, row * w1.cout_pad, cudaMemcpyDeviceToDevice));
        MDB_CUDA(cudaMemcpy(reinterpret_cast<char*>(cw.w) + row * w1.cout_pad, w2.w, row * w2.cout_pad, cudaMemcpyDeviceToDevice));
        MDB_CUDA(cudaMemcpy(cw.b, w1.b, sizeof(float) * w1.cout_pad, cudaMemcpyDeviceToDevice));
        MDB_CUDA(cudaMemcpy(cw.b + w1.cout_pad, w2.b, sizeof(float) * w2.cout_pad, cudaMemcpyDeviceToDevice));
        e->weights[b + ".cv12.conv"] = cw;
      }
    }
    e->plans.clear();  // plans hold weight pointers
    e->plan_lru.clear();
    e->last_plan = nullptr;
    for (int sl = 0; sl < 2; ++sl) { e->slot_plan[sl] = nullptr; e->slot_n[sl] = 0; }
    e->weights_loaded = true;
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

int mdb_infer_batch(mdb_engine* e, const uint8_t* const* rgb_hwc, const int32_t* h, const int32_t* w,
                    const int64_t* row_stride, int32_t n, float conf_thres, float* det_out, int32_t* det_count,
                    int32_t* kept_anchor, int32_t* status) {
  if (!e || !rgb_hwc || !h || !w || !det_out || !det_count || !status || n < 0) return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    std::vector<ImgIn> all(n);
    for (int i = 0; i < n; ++i)
      all[i] = ImgIn{rgb_hwc[i], h[i], w[i], row_stride ? row_stride[i] : static_cast<long long>(w[i]) * 3, h[i], w[i], i};
    return infer_common(e, all, false, conf_thres, det_out, det_count, kept_anchor, status, n);
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
}

int mdb_infer_preprocessed(mdb_engine* e, const uint8_t* const* lb_hwc, const int32_t* lh, const int32_t* lw,
                           const int32_t* src_h, const int32_t* src_w, int32_t n, float conf_thres, float* det_out,
                           int32_t* det_count, int32_t* kept_anchor, int32_t* status) {
  if (!e || !lb_hwc || !lh || !lw || !src_h || !src_w || !det_out || !det_count || !status || n < 0)
    return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    std::vector<ImgIn> all(n);
    for (int i = 0; i < n; ++i)
      all[i] = ImgIn{lb_hwc[i], lh[i], lw[i], static_cast<long long>(lw[i]) * 3, src_h[i], src_w[i], i};
    return infer_common(e, all, true, conf_thres, det_out, det_count, kept_anchor, status, n);
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
}

int mdb_infer_preprocessed_rp(mdb_engine* e, const uint8_t* const* lb_hwc, const int32_t* lh, const int32_t* lw,
                              const int32_t* src_h, const int32_t* src_w, const float* gain_pad, int32_t n, float conf_thres,
                              float* det_out, int32_t* det_count, int32_t* kept_anchor, int32_t* status) {
  if (!e || !lb_hwc || !lh || !lw || !src_h || !src_w || !det_out || !det_count || !status || n < 0)
    return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    std::vector<ImgIn> all(n);
    for (int i = 0; i < n; ++i) {
      all[i] = ImgIn{lb_hwc[i], lh[i], lw[i], static_cast<long long>(lw[i]) * 3, src_h[i], src_w[i], i};
      if (gain_pad) { all[i].rp_gain = gain_pad[3 * i]; all[i].rp_pad_x = gain_pad[3 * i + 1]; all[i].rp_pad_y = gain_pad[3 * i + 2]; }
    }
    return infer_common(e, all, true, conf_thres, det_out, det_count, kept_anchor, status, n);
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
}

int mdb_infer_jpeg_batch(mdb_engine* e, const uint8_t* const* jpeg, const size_t* nbytes, int32_t n, float conf_thres,
                         float* det_out, int32_t* det_count, int32_t* kept_anchor, int32_t* status, int32_t* src_hw) {
  if (!e || !jpeg || !nbytes || !det_out || !det_count || !status || n < 0) return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    if (!e->jpeg) e->jpeg = jpeg_create();
    std::vector<JpegImage> dec(n);
    e->last_jpeg_backend = jpeg_decode_batch(e->jpeg, jpeg, nbytes, n, dec.data(), e->st_copy);
    std::vector<ImgIn> all(n);
    for (int i = 0; i < n; ++i) {
      ImgIn im{dec[i].ptr, dec[i].h, dec[i].w, static_cast<long long>(dec[i].w) * 3, dec[i].h, dec[i].w, i};
      im.on_device = true;
      all[i] = im;
      if (src_hw) { src_hw[2 * i] = dec[i].h; src_hw[2 * i + 1] = dec[i].w; }
    }
    return infer_common(e, all, false, conf_thres, det_out, det_count, kept_anchor, status, n);
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
}

int mdb_op_decode_jpeg(mdb_engine* e, const uint8_t* jpeg, size_t nbytes, uint8_t* rgb_out, size_t cap, int32_t* hw2,
                       int32_t* backend) {
  if (!e || !jpeg || !hw2) return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    if (!e->jpeg) e->jpeg = jpeg_create();
    JpegImage im;
    const int be = jpeg_decode_batch(e->jpeg, &jpeg, &nbytes, 1, &im, e->st_copy);
    if (backend) *backend = be;
    hw2[0] = im.h; hw2[1] = im.w;
    if (!im.ptr) return fail(e, MDB_ERR_ARG, "not a decodable 1- or 3-component JPEG");
    if (rgb_out) {
      const size_t need = static_cast<size_t>(im.h) * im.w * 3;
      if (cap < need) return fail(e, MDB_ERR_ARG, "destination too small");
      MDB_CUDA(cudaStreamSynchronize(e->st_copy));
      MDB_CUDA(cudaMemcpy(rgb_out, im.ptr, need, cudaMemcpyDeviceToHost));
    }
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

int mdb_set_image_size(mdb_engine* e, int32_t img_size) {
  if (!e || img_size < 64 || img_size > 8192) return fail(e, MDB_ERR_ARG, "image size out of range");
  e->cfg.img_size = img_size;
  return MDB_OK;
}

int mdb_set_augment(mdb_engine* e, int32_t on) {
  if (!e) return MDB_ERR_ARG;
  e->augment = on != 0;
  if (e->augment) e->max_plans = std::max(e->max_plans, 7);  // three plans per shape group, two groups in flight
  return MDB_OK;
}

int mdb_set_pipeline_split(mdb_engine* e, int32_t parts) {
  if (!e || parts < 1 || parts > 8) return fail(e, MDB_ERR_ARG, "pipeline split out of range");
  e->pipeline_split = parts;
  return MDB_OK;
}

int mdb_bench_resident(mdb_engine* e, int32_t iters, float conf_thres, int32_t stages_mask, float* ms_total) {
  if (!e || iters < 1 || !ms_total) return fail(e, MDB_ERR_ARG, "bad argument");
  const int slot = e->last_slot;
  Plan* P = e->slot_plan[slot];
  if (!P || e->slot_n[slot] <= 0) return fail(e, MDB_ERR_STATE, "no batch is staged: call mdb_infer_batch first");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t st = e->st_compute;
    MDB_CUDA(cudaStreamSynchronize(st));
    MDB_CUDA(cudaEventRecord(e->ev_t[0], st));
    MDB_CUDA(cudaStreamSynchronize(e->st_post));
    for (int i = 0; i < iters; ++i) stage_compute(e, P, e->slot_n[slot], slot, conf_thres, stages_mask, false);
    MDB_CUDA(cudaEventRecord(e->ev_post_done, e->st_post));
    MDB_CUDA(cudaStreamWaitEvent(st, e->ev_post_done, 0));
    MDB_CUDA(cudaEventRecord(e->ev_t[1], st));
    MDB_CUDA(cudaEventSynchronize(e->ev_t[1]));
    MDB_CUDA(cudaEventElapsedTime(ms_total, e->ev_t[0], e->ev_t[1]));
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

int mdb_profile_steps(mdb_engine* e, int32_t iters, float* ms_out, double* macs_out, char* names_out,
                      int32_t name_stride, int32_t cap, int32_t* n_steps) {
  if (!e || iters < 1 || !n_steps) return fail(e, MDB_ERR_ARG, "bad argument");
  const int slot = e->last_slot;
  Plan* P = e->slot_plan[slot];
  if (!P || e->slot_n[slot] <= 0) return fail(e, MDB_ERR_STATE, "no batch is staged: call mdb_infer_batch first");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t st = e->st_compute;
    const int n = e->slot_n[slot];
    const int total = static_cast<int>(P->steps.size());
    *n_steps = total;
    for (int i = 0; i < total && i < cap; ++i) {
      Step& s = P->steps[i];
      MDB_CUDA(cudaStreamSynchronize(st));
      MDB_CUDA(cudaEventRecord(e->ev_t[0], st));
      for (int it = 0; it < iters; ++it) {
        if (s.kind == STEP_CONV) conv_launch(s.op, n, st);
        else if (s.kind == STEP_UPSAMPLE) upsample2x_launch(P->view(s.in), P->view(s.out), n, st);
        else {
          const VBuf& b = P->bufs[s.in.buf];
          const int hs = b.C / 4;
          View cat = P->view(s.in);
          View s1 = cat, p1 = cat, p2 = cat, p3 = cat;
          s1.c_off = 0; p1.c_off = hs; p2.c_off = 2 * hs; p3.c_off = 3 * hs;
          s1.C = p1.C = p2.C = p3.C = hs;
          sppf_pool_launch(s1, p1, p2, p3, n, st);
        }
      }
      MDB_CUDA(cudaEventRecord(e->ev_t[1], st));
      MDB_CUDA(cudaEventSynchronize(e->ev_t[1]));
      float ms = 0;
      MDB_CUDA(cudaEventElapsedTime(&ms, e->ev_t[0], e->ev_t[1]));
      e->launches += iters;
      if (ms_out) ms_out[i] = ms / iters;
      if (macs_out) macs_out[i] = s.kind == STEP_CONV ? conv_macs_per_image(s.op) * n : 0.0;
      if (names_out && name_stride > 1) {
        char buf[160];
        if (s.kind == STEP_CONV) {
          const VBuf& ib = P->bufs[s.in.buf];
          snprintf(buf, sizeof(buf), "%s k%d s%d %dx%d %d->%d", s.name.c_str(), s.k, s.s, ib.H, ib.W, s.in.C, s.cout);
        } else {
          snprintf(buf, sizeof(buf), "%s", s.name.c_str());
        }
        strncpy(names_out + static_cast<size_t>(i) * name_stride, buf, name_stride - 1);
        names_out[static_cast<size_t>(i) * name_stride + name_stride - 1] = 0;
      }
    }
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

int mdb_batch_begin(mdb_engine* e, const uint8_t* const* rgb_hwc, const int32_t* h, const int32_t* w,
                    const int64_t* row_stride, int32_t n, float conf_thres, float* det_out, int32_t* det_count,
                    int32_t* kept_anchor, int32_t* status) {
  if (!e || !rgb_hwc || !h || !w || !det_out || !det_count || !status || n < 0) return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    std::vector<ImgIn> all(n);
    for (int i = 0; i < n; ++i)
      all[i] = ImgIn{rgb_hwc[i], h[i], w[i], row_stride ? row_stride[i] : static_cast<long long>(w[i]) * 3, h[i], w[i], i};
    return batch_begin(e, all, false, conf_thres, det_out, det_count, kept_anchor, status);
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
}

int mdb_batch_next(mdb_engine* e, int32_t* done_idx, int32_t cap, int32_t* n_done) {
  if (!e || !done_idx || !n_done || cap < 1) return fail(e, MDB_ERR_ARG, "null argument");
  if (cudaSetDevice(e->cfg.device) != cudaSuccess) return fail(e, MDB_ERR_CUDA, "cudaSetDevice failed");
  return batch_next(e, done_idx, cap, n_done);
}

int mdb_last_timing(mdb_engine* e, float* ms4) {
  if (!e || !ms4) return MDB_ERR_ARG;
  memcpy(ms4, e->last_ms, sizeof(float) * 4);
  return MDB_OK;
}

int64_t mdb_launch_count(const mdb_engine* e) { return e ? e->launches : 0; }

int mdb_debug_fetch(mdb_engine* e, int32_t what, int32_t index, void* host_dst, size_t bytes, int64_t* shape4) {
  if (!e) return MDB_ERR_ARG;
  if (!e->last_plan) return fail(e, MDB_ERR_STATE, "no batch has been run yet");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    Plan* P = e->last_plan;
    const int n = e->last_n;
    const void* src = nullptr;
    size_t need = 0;
    int64_t shp[4] = {0, 0, 0, 0};
    if (what == MDB_FETCH_STEM_INPUT) {
      View v = P->view(VView{P->stem_buf, 0, 48});
      src = v.base; shp[0] = n; shp[1] = v.H; shp[2] = v.W; shp[3] = 48;
      need = static_cast<size_t>(n) * v.H * v.W * 48 * 2;
    } else if (what == MDB_FETCH_PRED) {
      const int no = e->cfg.arch.nc + 5;
      const bool tta = e->last_augment && P->pred_tta != nullptr;
      const int rows = tta ? P->anchors_tta : P->anchors;
      src = tta ? P->pred_tta : P->pred; shp[0] = n; shp[1] = rows; shp[2] = no; shp[3] = 1;
      need = static_cast<size_t>(n) * rows * no * 4;
    } else if (what == MDB_FETCH_LAYER) {
      int ci = -1;
      const Step* st = nullptr;
      for (auto& s : P->steps) {
        if (s.kind != STEP_CONV) continue;
        if (++ci == index) { st = &s; break; }
      }
      MDB_REQUIRE(st != nullptr, "layer index out of range");
      const VBuf& b = P->bufs[st->out.buf];
      // NB: returns the whole underlying buffer (all channels); valid only while its arena slot has not been
      // reused by a later layer -- the caller passes c_off/C through shape: {n*H*W, Ctot, c_off, C}
      src = static_cast<char*>(P->arena) + b.offset;
      shp[0] = static_cast<int64_t>(n) * b.H * b.W; shp[1] = b.C; shp[2] = st->out.c_off; shp[3] = st->out.C;
      need = static_cast<size_t>(n) * b.H * b.W * b.C * (b.fp32 ? 4 : 2);
    } else {
      return fail(e, MDB_ERR_ARG, "unknown fetch kind");
    }
    if (shape4) memcpy(shape4, shp, sizeof(shp));
    if (host_dst) {
      if (bytes < need) return fail(e, MDB_ERR_ARG, "destination too small");
      MDB_CUDA(cudaStreamSynchronize(e->st_compute));
      MDB_CUDA(cudaStreamSynchronize(e->st_post));
      MDB_CUDA(cudaMemcpy(host_dst, src, need, cudaMemcpyDeviceToHost));
    }
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}


int mdb_debug_set_trace(mdb_engine* e, const char* names) {
  if (!e) return MDB_ERR_ARG;
  e->trace_all = false;
  e->trace_names.clear();
  if (names == nullptr || names[0] == 0) return MDB_OK;
  const std::string s(names);
  if (s == "*") { e->trace_all = true; return MDB_OK; }
  size_t b = 0;
  while (b <= s.size()) {
    const size_t c = s.find(',', b);
    const std::string nm = s.substr(b, c == std::string::npos ? std::string::npos : c - b);
    if (!nm.empty()) e->trace_names.push_back(nm);
    if (c == std::string::npos) break;
    b = c + 1;
  }
  return MDB_OK;
}

int mdb_debug_halo_layout(int32_t mode, int32_t* out64) {
  if (!out64 || mode < 0 || mode > 2) return MDB_ERR_ARG;
  try {
    const HaloLayout L = halo_layout(mode);
    int32_t* o = out64;
    for (int i = 0; i < 64; ++i) o[i] = 0;
    o[0] = L.n_boxes; o[1] = static_cast<int32_t>(L.a_tx); o[2] = static_cast<int32_t>(L.a_stage_bytes);
    for (int b = 0; b < 4; ++b) {
      o[4 + b] = static_cast<int32_t>(L.box_off[b]); o[8 + b] = L.box_map[b]; o[12 + b] = L.box_par[b];
      o[16 + b] = L.box_dx[b]; o[20 + b] = L.box_dy[b]; o[24 + b] = L.box_w[b]; o[28 + b] = L.box_h[b];
    }
    for (int t = 0; t < 9; ++t) { o[32 + t] = static_cast<int32_t>(L.tap_off[t]); o[41 + t] = static_cast<int32_t>(L.tap_sbo[t]); }
  } catch (...) {
    return MDB_ERR_ARG;
  }
  return MDB_OK;
}

int mdb_debug_trace_names(mdb_engine* e, char* names_out, size_t cap, int32_t* n_steps) {
  if (!e || !n_steps) return MDB_ERR_ARG;
  if (!e->last_plan) return fail(e, MDB_ERR_STATE, "no batch has been run yet");
  std::string all;
  int n = 0;
  for (auto& s : e->last_plan->steps) {
    if (e->last_plan->trace.count(s.name) == 0) continue;
    all += s.name;
    all += '\n';
    ++n;
  }
  *n_steps = n;
  if (names_out && cap > 0) {
    strncpy(names_out, all.c_str(), cap - 1);
    names_out[cap - 1] = 0;
  }
  return MDB_OK;
}

int mdb_debug_fetch_trace(mdb_engine* e, const char* name, int32_t which, void* host_dst, size_t bytes, int64_t* shape5) {
  if (!e || !name || which < 0 || which > 2) return MDB_ERR_ARG;
  if (!e->last_plan) return fail(e, MDB_ERR_STATE, "no batch has been run yet");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    Plan* P = e->last_plan;
    auto it = P->trace.find(name);
    if (it == P->trace.end()) return fail(e, MDB_ERR_ARG, std::string("step was not traced: ") + name);
    const Plan::TraceRec& r = it->second;
    const bool f32 = which == 1 && r.out_fp32;
    int64_t shp[5] = {r.shape[which][0], r.shape[which][1], r.shape[which][2], r.shape[which][3], f32 ? 1 : 0};
    if (!r.present[which]) shp[0] = shp[1] = shp[2] = shp[3] = 0;
    if (shape5) memcpy(shape5, shp, sizeof(shp));
    if (host_dst && r.present[which]) {
      const size_t need = static_cast<size_t>(shp[0]) * shp[1] * shp[2] * shp[3] * (f32 ? 4 : 2);
      if (bytes < need) return fail(e, MDB_ERR_ARG, "destination too small");
      MDB_CUDA(cudaStreamSynchronize(e->st_compute));
      MDB_CUDA(cudaMemcpy(host_dst, static_cast<char*>(P->trace_mem) + r.off[which], need, cudaMemcpyDeviceToHost));
    }
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

// ---------------------------------------------------------------------------------------------
// single-kernel entry points (parity tests)
// ---------------------------------------------------------------------------------------------
int mdb_op_upsample2x(mdb_engine* e, const void* src_fp16, int32_t n, int32_t h, int32_t w, int32_t c, void* dst_fp16) {
  if (!e || !src_fp16 || !dst_fp16 || n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 != 0) return fail(e, MDB_ERR_ARG, "bad argument");
  void *d_s = nullptr, *d_d = nullptr;
  int rc = MDB_OK;
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    const size_t sb = static_cast<size_t>(n) * h * w * c * 2, db = sb * 4;
    MDB_CUDA(cudaMalloc(&d_s, sb));
    MDB_CUDA(cudaMalloc(&d_d, db));
    MDB_CUDA(cudaMemcpy(d_s, src_fp16, sb, cudaMemcpyHostToDevice));
    MDB_CUDA(cudaMemset(d_d, 0xFF, db));
    View s, d;
    s.base = d_s; s.N = n; s.H = h; s.W = w; s.Ctot = c; s.C = c;
    d.base = d_d; d.N = n; d.H = 2 * h; d.W = 2 * w; d.Ctot = c; d.C = c;
    upsample2x_launch(s, d, n, e->st_compute);
    e->launches++;
    MDB_CUDA(cudaStreamSynchronize(e->st_compute));
    MDB_CUDA(cudaMemcpy(dst_fp16, d_d, db, cudaMemcpyDeviceToHost));
  } catch (const CudaError& ex) {
    rc = fail(e, MDB_ERR_CUDA, ex.msg);
  }
  cudaFree(d_s); cudaFree(d_d);
  return rc;
}

int mdb_op_sppf_pool(mdb_engine* e, const void* s1_fp16, int32_t n, int32_t h, int32_t w, int32_t c, void* cat_fp16) {
  if (!e || !s1_fp16 || !cat_fp16 || n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 != 0) return fail(e, MDB_ERR_ARG, "bad argument");
  void* d_cat = nullptr;
  int rc = MDB_OK;
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    const size_t npix = static_cast<size_t>(n) * h * w;
    MDB_CUDA(cudaMalloc(&d_cat, npix * 4 * c * 2));
    MDB_CUDA(cudaMemset(d_cat, 0xFF, npix * 4 * c * 2));
    MDB_CUDA(cudaMemcpy2D(d_cat, static_cast<size_t>(4) * c * 2, s1_fp16, static_cast<size_t>(c) * 2, static_cast<size_t>(c) * 2, npix,
                          cudaMemcpyHostToDevice));
    View v[4];
    for (int i = 0; i < 4; ++i) { v[i].base = d_cat; v[i].N = n; v[i].H = h; v[i].W = w; v[i].Ctot = 4 * c; v[i].c_off = i * c; v[i].C = c; }
    sppf_pool_launch(v[0], v[1], v[2], v[3], n, e->st_compute);
    e->launches++;
    MDB_CUDA(cudaStreamSynchronize(e->st_compute));
    MDB_CUDA(cudaMemcpy(cat_fp16, d_cat, npix * 4 * c * 2, cudaMemcpyDeviceToHost));
  } catch (const CudaError& ex) {
    rc = fail(e, MDB_ERR_CUDA, ex.msg);
  }
  cudaFree(d_cat);
  return rc;
}

int mdb_op_decode(mdb_engine* e, const float* const* raw4, const int32_t* ny4, const int32_t* nx4, int32_t n, float* pred) {
  if (!e || !raw4 || !ny4 || !nx4 || !pred || n <= 0) return fail(e, MDB_ERR_ARG, "bad argument");
  float* d_raw[4] = {nullptr, nullptr, nullptr, nullptr};
  float* d_pred = nullptr;
  int rc = MDB_OK;
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    const int no = e->cfg.arch.nc + 5;
    const int craw = ((3 * no + 3) / 4) * 4;
    const float strides[4] = {8.f, 16.f, 32.f, 64.f};
    DecodeLevel lv[4];
    int off = 0;
    for (int l = 0; l < 4; ++l) {
      const size_t b = static_cast<size_t>(n) * ny4[l] * nx4[l] * craw * 4;
      MDB_REQUIRE(raw4[l] != nullptr && ny4[l] > 0 && nx4[l] > 0, "bad level");
      MDB_CUDA(cudaMalloc(&d_raw[l], b));
      MDB_CUDA(cudaMemcpy(d_raw[l], raw4[l], b, cudaMemcpyHostToDevice));
      lv[l].raw = d_raw[l]; lv[l].ny = ny4[l]; lv[l].nx = nx4[l]; lv[l].c_stride = craw; lv[l].stride = strides[l]; lv[l].row_off = off;
      for (int a = 0; a < 3; ++a) { lv[l].anchors[a][0] = e->cfg.arch.anchors_px[l][a][0]; lv[l].anchors[a][1] = e->cfg.arch.anchors_px[l][a][1]; }
      off += 3 * ny4[l] * nx4[l];
    }
    const size_t pb = static_cast<size_t>(n) * off * no * 4;
    MDB_CUDA(cudaMalloc(&d_pred, pb));
    decode_launch(lv, n, no, off, d_pred, e->st_compute);
    e->launches++;
    MDB_CUDA(cudaStreamSynchronize(e->st_compute));
    MDB_CUDA(cudaMemcpy(pred, d_pred, pb, cudaMemcpyDeviceToHost));
  } catch (const CudaError& ex) {
    rc = fail(e, MDB_ERR_CUDA, ex.msg);
  }
  for (int l = 0; l < 4; ++l) cudaFree(d_raw[l]);
  cudaFree(d_pred);
  return rc;
}

int mdb_op_letterbox(mdb_engine* e, const uint8_t* rgb_hwc, int32_t h, int32_t w, int64_t row_stride, void* dst_fp16,
                     int32_t* out_hw2) {
  if (!e || !rgb_hwc || !dst_fp16 || h <= 0 || w <= 0) return fail(e, MDB_ERR_ARG, "null argument");
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    const mdb_config& cfg = e->cfg;
    const LetterboxGeom g = geometry_for(cfg, h, w);
    MDB_REQUIRE(g.out_h % 2 == 0 && g.out_w % 2 == 0, "letterboxed size must be even");
    cudaStream_t st = e->st_compute;
    const size_t row = static_cast<size_t>(w) * 3;
    uint8_t* d_src = nullptr;
    __half* d_dst = nullptr;
    PreImage* d_pi = nullptr;
    const size_t out_elems = static_cast<size_t>(g.out_h / 2) * (g.out_w / 2) * 48;
    MDB_CUDA(cudaMalloc(&d_src, row * h + 256));
    MDB_CUDA(cudaMalloc(&d_dst, out_elems * 2));
    MDB_CUDA(cudaMalloc(&d_pi, sizeof(PreImage)));
    MDB_CUDA(cudaMemcpy2D(d_src, row, rgb_hwc, row_stride ? static_cast<size_t>(row_stride) : row, row, h, cudaMemcpyHostToDevice));
    PreImage pi;
    pi.src = d_src; pi.src_stride = static_cast<long long>(row); pi.src_h = h; pi.src_w = w;
    pi.new_w = g.new_w; pi.new_h = g.new_h; pi.top = g.top; pi.left = g.left;
    fill_resize(e, pi, h, w, g);
    MDB_CUDA(cudaMemcpy(d_pi, &pi, sizeof(pi), cudaMemcpyHostToDevice));
    const float max_scale = std::max(1.0f, std::max(static_cast<float>(w) / g.new_w, static_cast<float>(h) / g.new_h));
    preprocess_launch(d_pi, 1, g.out_h, g.out_w, d_dst, max_scale, st);
    e->launches++;
    MDB_CUDA(cudaStreamSynchronize(st));
    MDB_CUDA(cudaMemcpy(dst_fp16, d_dst, out_elems * 2, cudaMemcpyDeviceToHost));
    if (out_hw2) { out_hw2[0] = g.out_h; out_hw2[1] = g.out_w; }
    cudaFree(d_src); cudaFree(d_dst); cudaFree(d_pi);
  } catch (const CudaError& ex) {
    return fail(e, MDB_ERR_CUDA, ex.msg);
  }
  return MDB_OK;
}

int mdb_op_conv2d(mdb_engine* e, const void* x_fp16, int32_t n, int32_t h, int32_t w, int32_t cin, const float* wgt,
                  const float* bias, int32_t cout, int32_t k, int32_t s, int32_t pad, int32_t act,
                  const void* residual_fp16, int32_t out_fp32, void* y, int32_t iters, float* ms) {
  if (!e || !x_fp16 || !wgt || !bias || !y) return fail(e, MDB_ERR_ARG, "null argument");
  void *d_x = nullptr, *d_y = nullptr, *d_r = nullptr;
  __half* d_w = nullptr;
  float* d_b = nullptr;
  ConvOp* op = nullptr;
  int rc = MDB_OK;
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    MDB_REQUIRE(cin % 16 == 0, "cin must be a multiple of 16");
    cudaStream_t st = e->st_compute;
    const int ho = (h + 2 * pad - k) / s + 1, wo = (w + 2 * pad - k) / s + 1;
    const int cout_pad = ((cout + 15) / 16) * 16;
    const int cstore = out_fp32 ? ((cout + 3) / 4) * 4 : ((cout + 7) / 8) * 8;
    MDB_REQUIRE(cstore == cout, "cout must be a multiple of 8 (fp16) / 4 (fp32) for this test entry point");
    std::vector<__half> packed;
    const int cin_pad = conv_cin_pad(cin);
    pack_weights_khwc(wgt, cout, cin, k, cout_pad, cin_pad, packed);
    std::vector<float> bp(cout_pad, 0.f);
    memcpy(bp.data(), bias, sizeof(float) * cout);
    const size_t xb = static_cast<size_t>(n) * h * w * cin * 2;
    const size_t yb = static_cast<size_t>(n) * ho * wo * cout * (out_fp32 ? 4 : 2);
    MDB_CUDA(cudaMalloc(&d_x, xb + 256));
    MDB_CUDA(cudaMalloc(&d_y, yb + 256));
    MDB_CUDA(cudaMalloc(&d_w, packed.size() * 2 + 256));
    MDB_CUDA(cudaMalloc(&d_b, bp.size() * 4));
    MDB_CUDA(cudaMemcpy(d_x, x_fp16, xb, cudaMemcpyHostToDevice));
    MDB_CUDA(cudaMemcpy(d_w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    MDB_CUDA(cudaMemcpy(d_b, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
    MDB_CUDA(cudaMemset(d_y, 0xFF, yb));
    ConvDesc d;
    d.in.base = d_x; d.in.N = n; d.in.H = h; d.in.W = w; d.in.Ctot = cin; d.in.C = cin;
    d.out.base = d_y; d.out.N = n; d.out.H = ho; d.out.W = wo; d.out.Ctot = cout; d.out.C = cout; d.out.fp32 = out_fp32 != 0;
    if (residual_fp16) {
      const size_t rb = static_cast<size_t>(n) * ho * wo * cout * 2;
      MDB_CUDA(cudaMalloc(&d_r, rb + 256));
      MDB_CUDA(cudaMemcpy(d_r, residual_fp16, rb, cudaMemcpyHostToDevice));
      d.res.base = d_r; d.res.N = n; d.res.H = ho; d.res.W = wo; d.res.Ctot = cout; d.res.C = cout;
    }
    d.wgt = d_w; d.bias = d_b; d.cout = cout; d.cout_pad = cout_pad; d.cin_pad = cin_pad; d.k = k; d.s = s; d.pad = pad; d.act = act;
    op = conv_build(d, e->num_sms);
    conv_launch(op, n, st);
    e->launches++;
    MDB_CUDA(cudaStreamSynchronize(st));
    if (iters > 1) {
      cudaEvent_t a, b;
      MDB_CUDA(cudaEventCreate(&a));
      MDB_CUDA(cudaEventCreate(&b));
      MDB_CUDA(cudaEventRecord(a, st));
      for (int i = 0; i < iters; ++i) conv_launch(op, n, st);
      MDB_CUDA(cudaEventRecord(b, st));
      MDB_CUDA(cudaEventSynchronize(b));
      float t = 0;
      cudaEventElapsedTime(&t, a, b);
      if (ms) *ms = t / iters;
      e->launches += iters;
      cudaEventDestroy(a); cudaEventDestroy(b);
    } else if (ms) {
      *ms = 0;
    }
    MDB_CUDA(cudaMemcpy(y, d_y, yb, cudaMemcpyDeviceToHost));
  } catch (const CudaError& ex) {
    rc = fail(e, MDB_ERR_CUDA, ex.msg);
  }
  if (op) conv_free(op);
  cudaFree(d_x); cudaFree(d_y); cudaFree(d_r); cudaFree(d_w); cudaFree(d_b);
  return rc;
}

int mdb_op_nms(mdb_engine* e, const float* pred, int32_t n, int32_t anchors, int32_t nc, float conf_thres,
               double iou_thres, int32_t max_det, float* det_out, int32_t* det_count, int32_t* kept_anchor) {
  if (!e || !pred || !det_out || !det_count || n <= 0 || anchors <= 0) return fail(e, MDB_ERR_ARG, "null argument");
  float* d_pred = nullptr; void* d_scr = nullptr; float* d_det = nullptr; int32_t* d_cnt = nullptr; int32_t* d_anc = nullptr;
  int rc = MDB_OK;
  try {
    MDB_CUDA(cudaSetDevice(e->cfg.device));
    MDB_REQUIRE(max_det >= 1 && max_det <= kMaxDetCap, "max_det out of range");
    cudaStream_t st = e->st_compute;
    const int no = nc + 5;
    const size_t pb = static_cast<size_t>(n) * anchors * no * 4;
    MDB_CUDA(cudaMalloc(&d_pred, pb));
    MDB_CUDA(cudaMalloc(&d_scr, nms_scratch_bytes(n, anchors)));
    MDB_CUDA(cudaMalloc(&d_det, static_cast<size_t>(n) * max_det * 6 * 4));
    MDB_CUDA(cudaMalloc(&d_cnt, n * 4));
    MDB_CUDA(cudaMalloc(&d_anc, static_cast<size_t>(n) * max_det * 4));
    MDB_CUDA(cudaMemcpy(d_pred, pred, pb, cudaMemcpyHostToDevice));
    MDB_CUDA(cudaMemset(d_det, 0, static_cast<size_t>(n) * max_det * 6 * 4));
    MDB_CUDA(cudaMemset(d_anc, 0xFF, static_cast<size_t>(n) * max_det * 4));
    NmsScratch sc;
    nms_bind_scratch(&sc, d_scr, n, anchors);
    nms_candidates_launch(d_pred, n, anchors, no, conf_thres, sc, st);
    e->launches++;
    nms_launch(d_pred, n, anchors, no, conf_thres, float_floor_of_double(iou_thres), max_det, nullptr, sc, d_det, d_cnt,
               d_anc, st);
    e->launches++;
    MDB_CUDA(cudaStreamSynchronize(st));
    MDB_CUDA(cudaMemcpy(det_out, d_det, static_cast<size_t>(n) * max_det * 6 * 4, cudaMemcpyDeviceToHost));
    MDB_CUDA(cudaMemcpy(det_count, d_cnt, n * 4, cudaMemcpyDeviceToHost));
    if (kept_anchor) MDB_CUDA(cudaMemcpy(kept_anchor, d_anc, static_cast<size_t>(n) * max_det * 4, cudaMemcpyDeviceToHost));
  } catch (const CudaError& ex) {
    rc = fail(e, MDB_ERR_CUDA, ex.msg);
  }
  cudaFree(d_pred); cudaFree(d_scr); cudaFree(d_det); cudaFree(d_cnt); cudaFree(d_anc);
  return rc;
}

}  // extern "C"
