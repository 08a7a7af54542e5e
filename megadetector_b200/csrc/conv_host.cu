// Host side of K2: tile-shape selection, TMA descriptors, launch.
#include <algorithm>
#include <cmath>

#include "common.h"
#include <cstdlib>
#include <vector>

#include "conv_pair_tcgen05.cuh"

namespace mdb {

struct ConvOp {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  ConvMaps maps;
  PairParams pp;
  bool flat = false;
  int HW_out = 0;  // output pixels per image
  size_t smem = 0;
  int num_sms = 148;
  double macs_per_image = 0;
};

int conv_pick_bn(int cout_pad) {
  // largest multiple of 16 that divides cout_pad and fits one TMEM accumulator stage (256 columns)
  for (int bn = 256; bn >= 16; bn -= 16)
    if (cout_pad % bn == 0) return bn;
  return 16;
}

int conv_cin_pad(int cin) {
  // Whole 64-channel chunks (128-byte swizzle rows): channels beyond the tensor's are zero-filled by the TMA unit (no memory
  // traffic) and meet zero weights.  Short rows make the TMA unit request-bound (3x3 80->80 ran 15x off the MMA floor with
  // 32-byte rows); narrow 16 / 32-channel tail chunks were implemented and measured in round 1 and cost as much
  // shared-memory fill time as a zero-filled 64-channel chunk (the TMA unit is bound by rows, not bytes:
  // profiles/r01_conv_tail_probe.txt, r01_conv_oob_probe.txt), so they were removed again.
  return ((cin + 63) / 64) * 64;
}

HaloLayout halo_layout(int mode) {
  // mode 0: stride 1, one 8 px x 18 row box per kw shift (the tap's kh shift is a whole 1024-byte image row of its box);
  // mode 1: stride 1, ONE 10 px x 18 row box: tap (kh, kw) starts kh rows (1280 bytes each) and kw pixels (128 bytes) in;
  // mode 2: stride 2, one box per parity view (py, px) = in[2r + py][2c + px]: output (y, x), tap (kh, kw) reads
  //         in[2y + kh - 1][2x + kw - 1], i.e. kh = 0 -> row y - 1 of the odd rows, kh = 1 -> row y of the even rows,
  //         kh = 2 -> row y of the odd rows (columns alike): boxes of 8|9 px x 16|17 rows starting one row / pixel early
  //         on the odd views.
  // Pure host arithmetic (tests/test_capi_host.py models the addressing from these tables without a GPU).
  HaloLayout L{};
  auto align1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  if (mode == 0) {
    L.n_boxes = 3;
    for (int b = 0; b < 3; ++b) {
      L.box_off[b] = static_cast<uint32_t>(b) * kHaloBox; L.box_map[b] = 0; L.box_par[b] = -1;
      L.box_dx[b] = b - 1; L.box_dy[b] = -1; L.box_w[b] = kHaloTw; L.box_h[b] = kHaloRows;
    }
    L.a_tx = 3 * kHaloBox;
    L.a_stage_bytes = 3 * kHaloBox;
    for (int t = 0; t < 9; ++t) {
      L.tap_off[t] = static_cast<uint32_t>(t % 3) * kHaloBox + static_cast<uint32_t>(t / 3) * 1024u;
      L.tap_sbo[t] = 1024u >> 4;
    }
  } else if (mode == 1) {
    L.n_boxes = 1;
    L.box_off[0] = 0; L.box_map[0] = 0; L.box_par[0] = -1; L.box_dx[0] = -1; L.box_dy[0] = -1;
    L.box_w[0] = kHaloTw + 2; L.box_h[0] = kHaloRows;
    L.a_tx = kHalo1Box;
    L.a_stage_bytes = kHalo1Stage;
    for (int t = 0; t < 9; ++t) {
      L.tap_off[t] = static_cast<uint32_t>(t % 3) * 128u + static_cast<uint32_t>(t / 3) * kHalo1Pitch;
      L.tap_sbo[t] = kHalo1Pitch >> 4;
    }
  } else {
    MDB_REQUIRE(mode == 2, "halo layout mode must be 0, 1 or 2");
    L.n_boxes = 4;
    uint32_t off = 0, pitch[4];
    for (int v = 0; v < 4; ++v) {
      const int py = v >> 1, px = v & 1;
      L.box_off[v] = off; L.box_map[v] = v; L.box_par[v] = v; L.box_dx[v] = -px; L.box_dy[v] = -py;
      L.box_w[v] = kHaloTw + px; L.box_h[v] = kHaloTh + py;
      pitch[v] = static_cast<uint32_t>(L.box_w[v]) * 128u;
      L.a_tx += static_cast<uint32_t>(L.box_w[v] * L.box_h[v]) * 128u;
      off += align1k(static_cast<uint32_t>(L.box_w[v] * L.box_h[v]) * 128u);
    }
    L.a_stage_bytes = off;
    for (int t = 0; t < 9; ++t) {
      const int kh = t / 3, kx = t % 3;
      const int v = (kh != 1 ? 2 : 0) | (kx != 1 ? 1 : 0);
      L.tap_off[t] = L.box_off[v] + (kh == 2 ? pitch[v] : 0u) + (kx == 2 ? 128u : 0u);
      L.tap_sbo[t] = pitch[v] >> 4;
    }
  }
  return L;
}

static void pick_tile(int H, int W, int* tw, int* th) {
  // maximise useful rows per 128-row MMA tile; prefer wide tiles (longer contiguous TMA rows)
  double best = -1;
  int bw = 1, bh = 1;
  for (int w = std::min(W, 128); w >= 1; --w) {
    int h = std::min(H, 128 / w);
    if (h < 1) continue;
    if (w > 256 || h > 256) continue;
    const long tiles = static_cast<long>((W + w - 1) / w) * ((H + h - 1) / h);
    const double eff = static_cast<double>(H) * W / (static_cast<double>(tiles) * 128.0);
    if (eff > best + 1e-9) {
      best = eff;
      bw = w;
      bh = h;
    }
  }
  *tw = bw;
  *th = bh;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: every device an engine is created on opts the
// kernel into the full 227 KB once (mdb_create -> conv_init_device), instead of a process-global "largest size seen" cache
// that would leave a second GPU of the same process at the 48 KB default.
void conv_init_device() {
  MDB_CUDA(cudaFuncSetAttribute(conv_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8192));
}

ConvOp* conv_build(const ConvDesc& d, int num_sms) {
  ConvOp* op = new ConvOp();
  try {
    const View& in = d.in;
    const View& out = d.out;
    constexpr int kc = 64;  // K chunk = one 128-byte swizzle row of channels (see conv_cin_pad)
    MDB_REQUIRE(in.C % 8 == 0, "conv input channels must be a multiple of 8");
    MDB_REQUIRE(d.cin_pad >= in.C && d.cin_pad % kc == 0, "C_in per tap must be padded to a multiple of 64");
    MDB_REQUIRE(d.cout_pad % 16 == 0 && d.cout_pad >= d.cout, "bad cout_pad");
    MDB_REQUIRE(d.s == 1 || d.s == 2, "stride must be 1 or 2");
    MDB_REQUIRE(out.C >= d.cout, "output view too narrow");
    op->num_sms = num_sms;
    const int kw = d.kw > 0 ? d.kw : d.k;
    const int pad_x = d.pad_x >= 0 ? d.pad_x : d.pad;
    const int Ho = (in.H + 2 * d.pad - d.k) / d.s + 1;
    const int Wo = (in.W + 2 * pad_x - kw) / d.s + 1;
    MDB_REQUIRE(Ho == out.H && Wo == out.W && out.N == in.N, "conv output view has the wrong shape");
    op->HW_out = Ho * Wo;
    op->macs_per_image = static_cast<double>(Ho) * Wo * d.cout * in.C * d.k * kw;
    op->flat = (d.k == 1 && d.s == 1 && d.pad == 0);
    if (d.res.base) {
      MDB_REQUIRE(d.res.H == Ho && d.res.W == Wo && d.res.N == in.N && !d.res.fp32, "residual view mismatch");
      MDB_REQUIRE(d.res.Ctot % 8 == 0 && d.res.c_off % 8 == 0, "residual view must be 16-byte aligned per pixel");
    }
    MDB_REQUIRE(out.fp32 ? (out.Ctot % 4 == 0 && out.c_off % 4 == 0) : (out.Ctot % 8 == 0 && out.c_off % 8 == 0),
                "output view must be 16-byte aligned per pixel");

    PairParams& q = op->pp;
    q.H_out = Ho; q.W_out = Wo;
    q.BN = conv_pick_bn(d.cout_pad);
    q.n_tiles = d.cout_pad / q.BN;
    q.Cout = d.cout;
    q.taps = d.k * kw;
    q.chunks = d.cin_pad / kc;
    q.total_chunks = q.taps * q.chunks;
    {
      // channels beyond in.C are TMA zero fill meeting zero weights: the MMA steps over them are skipped (exact: they add 0).
      // C_in = 80 runs 5 instead of 8 steps per tap, 160: 10 of 12, 480: 30 of 32, the stem's 48: 3 of 4.  MDB_CONV_TAILK=0: off.
      static const bool tailk = !(getenv("MDB_CONV_TAILK") && getenv("MDB_CONV_TAILK")[0] == '0');
      const int rest = in.C - (q.chunks - 1) * kc;
      q.tail_k = tailk ? std::min(kc / 16, std::max(1, (rest + 15) / 16)) : kc / 16;
    }
    MDB_REQUIRE(q.taps <= 9, "at most 9 filter taps");
    MDB_REQUIRE(d.cout_pad <= kPairMaxCout, "at most 1280 (padded) output channels");
    if (op->flat) {
      q.tw = 128;
      q.th = 1;
      make_act_tmap_flat(&op->tmA[0], in, kc);
      op->tmA[1] = op->tmA[2] = op->tmA[3] = op->tmA[0];
    } else {
      pick_tile(Ho, Wo, &q.tw, &q.th);
      if (d.s == 1) {
        make_act_tmap(&op->tmA[0], in, kc, q.tw, q.th, -1, -1);
        op->tmA[1] = op->tmA[2] = op->tmA[3] = op->tmA[0];
      } else {
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) make_act_tmap(&op->tmA[py * 2 + px], in, kc, q.tw, q.th, py, px);
      }
    }
    // CTA pair: each CTA stages half of the B tile
    make_wgt_tmap(&op->tmB, d.wgt, d.k * kw * d.cin_pad, d.cout_pad, kc, q.BN / 2);
    for (int i = 0; i < 10; ++i) q.tap_word[i] = 0;
    for (int kh = 0; kh < d.k; ++kh)
      for (int kx = 0; kx < kw; ++kx) {
        int dx, dy, par = 0;
        if (d.s == 1) {
          dx = kx - pad_x;
          dy = kh - d.pad;
        } else {  // the tap reads one of the four parity views of the input, at half resolution
          const int ty = kh - d.pad, tx = kx - pad_x;
          const int py = ty & 1, px = tx & 1;
          dy = (ty - py) >> 1;
          dx = (tx - px) >> 1;
          par = py * 2 + px;
        }
        MDB_REQUIRE(dx >= -8 && dx < 120 && dy >= -8 && dy < 120, "filter tap offset out of range");
        q.tap_word[kh * kw + kx] = static_cast<uint32_t>(dx + 8) | static_cast<uint32_t>(dy + 8) << 8 | static_cast<uint32_t>(par) << 16;
      }
    // chunks per pipeline stage: enough MMA time per mbarrier round trip (>= ~320 tensor cycles)
    {
      static const char* e = getenv("MDB_CONV_CPS");
      q.cps = e ? atoi(e) : 2;  // measured sweep (profiles/r01_conv_cases_cps*.txt): 2 wins on every layer shape
      if (q.cps < 1) q.cps = 1;
      if (q.cps > 4) q.cps = 4;
    }
    q.bias = d.bias;
    q.out = out.base;
    q.out_pix_stride = out.Ctot;
    q.out_c_off = out.c_off;
    q.out_fp32 = out.fp32 ? 1 : 0;
    q.res = d.res.base ? static_cast<const __half*>(d.res.base) : nullptr;
    q.res_pix_stride = d.res.Ctot;
    q.res_c_off = d.res.c_off;
    q.act = d.act;
    q.idesc = make_idesc_f16(256, static_cast<uint32_t>(q.BN));
    q.dbg = nullptr;
    {
      // Epilogue-bound layers (MMA time per tile = K/64 x 16 x BN/... < epilogue time ~ 14 x BN cycles, i.e. total K <~ 1024):
      // the blocked producer / MMA threads back off instead of spinning (MDB_CONV_BACKOFF=<ns>, 0 = spin)
      static const char* eb = getenv("MDB_CONV_BACKOFF");
      const int ns = eb ? atoi(eb) : 0;
      q.backoff_ns = (q.total_chunks * kc <= 1024) ? ns : 0;
    }
    q.halo = 0;
    q.n_boxes = 0; q.a_stage_bytes = 0; q.a_tx = 0;
    for (int i = 0; i < 4; ++i) { q.box_off[i] = 0; q.box_map[i] = 0; q.box_dx[i] = 0; q.box_dy[i] = 0; }
    for (int i = 0; i < 9; ++i) { q.tap_off[i] = 0; q.tap_sbo[i] = 0; }
    q.a_stages = 0;
    {
      static const char* ep = getenv("MDB_CONV_PROBE");  // honoured by -DMDB_CONV_PROBES builds only
      q.probe = ep ? atoi(ep) : 0;
    }
    {
      const int oa = out.fp32 ? 8 : 16;  // elements per 32 bytes
      bool wide = (out.Ctot % oa == 0) && (out.c_off % oa == 0) && (reinterpret_cast<uintptr_t>(out.base) % 32 == 0);
      if (d.res.base)
        wide = wide && (d.res.Ctot % 16 == 0) && (d.res.c_off % 16 == 0) && (reinterpret_cast<uintptr_t>(d.res.base) % 32 == 0);
      q.wide = wide ? 1 : 0;
      q.fast = (wide && !out.fp32 && d.cout % 16 == 0 && d.cout_pad == d.cout) ? 1 : 0;
    }
    const size_t a_chunk = static_cast<size_t>(kTileM) * kc * 2;
    const size_t b_chunk = (static_cast<size_t>(q.BN / 2) * kc * 2 + 1023) & ~static_cast<size_t>(1023);
    // 3x3 layers (stride 1 or 2) whose output tiles well with 8 x 16 blocks use the halo main loop
    {
      static const char* eh = getenv("MDB_CONV_HALO");
      const bool allow = !(eh && eh[0] == '0');
      static const char* ehm = getenv("MDB_CONV_HALO_MIN");
      const int halo_min_chunks = ehm ? atoi(ehm) : 2;
      // MDB_CONV_HALO1=0: stride 1 fetches three 8-pixel boxes (one per kw shift) instead of one 10-pixel box
      static const char* eh1 = getenv("MDB_CONV_HALO1");
      const bool single = !(eh1 && eh1[0] == '0');
      // MDB_CONV_HALO_S2=0: stride 2 keeps the per-tap main loop
      static const char* eh2 = getenv("MDB_CONV_HALO_S2");
      const bool allow_s2 = !(eh2 && eh2[0] == '0');
      static const char* eas = getenv("MDB_CONV_HALO_ASTAGES");
      const double eff = static_cast<double>(Ho) * Wo / (static_cast<double>((Wo + 7) / 8 * 8) * ((Ho + 15) / 16 * 16));
      const bool shape_ok = allow && d.k == 3 && kw == 3 && d.pad == 1 && pad_x == 1 && eff >= 0.85 && q.chunks >= halo_min_chunks;
      const size_t budget = 227 * 1024 - 8192 - 1024;  // static: barriers + the bias vector; alignment slack
      bool use = false;
      HaloLayout L{};
      if (shape_ok && d.s == 1) {
        // measured (one filter row per weight stage): -26% time at 80 ch, -16% at 160, -9% at 320; single box: another -2..-6 %
        use = true;
        L = halo_layout(single ? 1 : 0);
        q.a_stages = eas ? std::max(2, std::min(kHaloMaxA, atoi(eas))) : (single ? 3 : 2);
      } else if (shape_ok && d.s == 2 && allow_s2) {
        L = halo_layout(2);
        q.a_stages = 2;
        use = 2 * static_cast<size_t>(L.a_stage_bytes) + 2 * 3 * b_chunk <= budget;  // two A stages + two weight stages must fit
        if (!use) q.a_stages = 0;
      }
      if (use) {
        q.n_boxes = L.n_boxes;
        q.a_tx = L.a_tx;
        q.a_stage_bytes = L.a_stage_bytes;
        for (int b = 0; b < L.n_boxes; ++b) {
          q.box_off[b] = L.box_off[b]; q.box_map[b] = L.box_map[b]; q.box_dx[b] = L.box_dx[b]; q.box_dy[b] = L.box_dy[b];
          const int par = L.box_par[b];
          make_act_tmap(&op->tmA[L.box_map[b]], in, kc, L.box_w[b], L.box_h[b], par < 0 ? -1 : par >> 1, par < 0 ? -1 : par & 1);
        }
        if (L.box_par[0] < 0)
          for (int i = 1; i < 4; ++i) op->tmA[i] = op->tmA[0];
        for (int t = 0; t < 9; ++t) { q.tap_off[t] = L.tap_off[t]; q.tap_sbo[t] = L.tap_sbo[t]; }
      }
      if (use) {
        q.halo = 1;
        q.tw = kHaloTw;
        q.th = kHaloTh;
        for (int i = 0; i < 4; ++i) op->maps.a[i] = op->tmA[i];
        op->maps.b = op->tmB;
        const size_t a_bytes = static_cast<size_t>(q.a_stages) * q.a_stage_bytes;
        int stages = static_cast<int>(std::min<size_t>(kPairMaxStages, (budget - a_bytes) / (3 * b_chunk)));  // a stage = one filter row
        MDB_REQUIRE(stages >= 2, "not enough shared memory for the halo pipeline");
        q.stages = stages;
        q.cps = 1;
        op->smem = a_bytes + 3 * b_chunk * stages + 1024;
        return op;
      }
    }
    const size_t stage = static_cast<size_t>(q.cps) * (a_chunk + b_chunk);
    const size_t budget = 227 * 1024 - 8192;  // static: barriers + the bias vector; alignment slack
    const int stage_iters = (q.total_chunks + q.cps - 1) / q.cps;
    int stages = static_cast<int>(std::min<size_t>(kPairMaxStages, (budget - 1024) / stage));
    stages = std::max(2, std::min(stages, std::max(2, stage_iters * 2)));
    q.stages = stages;
    op->smem = stage * stages + 1024;
    for (int i = 0; i < 4; ++i) op->maps.a[i] = op->tmA[i];
    op->maps.b = op->tmB;
  } catch (...) {
    delete op;
    throw;
  }
  return op;
}

static void conv_launch_pair(ConvOp* op, int n_img, cudaStream_t st) {
  PairParams p = op->pp;
  if (op->flat) {
    const long long npix = static_cast<long long>(n_img) * op->HW_out;
    p.n_img = 1;
    p.H_out = 1;
    p.W_out = static_cast<int>(npix);
    p.tiles_x = static_cast<int>((npix + 127) / 128);
    p.tiles_y = 1;
  } else {
    p.n_img = n_img;
    p.tiles_x = (p.W_out + p.tw - 1) / p.tw;
    p.tiles_y = (p.H_out + p.th - 1) / p.th;
  }
  p.m_tiles = p.n_img * p.tiles_x * p.tiles_y;
  p.pair_m = (p.m_tiles + 1) / 2;
  const long long total = static_cast<long long>(p.pair_m) * p.n_tiles;
  if (total <= 0) return;
  // division-free tile decode: floor(x * ceil(2^32 / d) / 2^32) == x / d for every x with x * d < 2^32
  auto magic = [&](int d, long long max_x) -> uint32_t {
    MDB_REQUIRE(d >= 1 && max_x * d < (1ll << 32), "tile count too large for the 32-bit division-free decode");
    return d == 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + static_cast<uint64_t>(d) - 1) / static_cast<uint64_t>(d));
  };
  p.mg_ntiles = magic(p.n_tiles, total);
  p.mg_tx = magic(p.tiles_x, 2ll * p.pair_m + 1);
  p.mg_ty = magic(p.tiles_y, 2ll * p.pair_m + 1);
  const int clusters = static_cast<int>(std::min<long long>(total, op->num_sms / 2));
  const int grid = clusters * 2;
  static const bool debug = getenv("MDB_CONV_DEBUG") && getenv("MDB_CONV_DEBUG")[0] == '1';
  long long* d_dbg = nullptr;
  if (debug) {
    MDB_CUDA(cudaMalloc(&d_dbg, sizeof(long long) * 16 * grid));
    MDB_CUDA(cudaMemsetAsync(d_dbg, 0, sizeof(long long) * 16 * grid, st));
  }
  p.dbg = d_dbg;
  // programmatic dependent launch: the next layer's CTAs start (barrier init, TMEM allocation, bias) as SMs free up and wait
  // for this layer's completion before touching activations.  -3 % network time at 8 images, neutral at 64; MDB_CONV_PDL=0: off
  static const bool pdl_on = !(getenv("MDB_CONV_PDL") && getenv("MDB_CONV_PDL")[0] == '0');
  // ... for launches of up to 16 images (MDB_CONV_PDL_MAX_IMAGES): the pipelined end-to-end path runs 16-image chunks and gains
  // (e2e 1486 vs 1472 images/s), the 64-image resident pass loses 0.6 % with it (profiles/r02_conv_ab_experiments.txt par. 14, 18)
  static const int pdl_max_images = getenv("MDB_CONV_PDL_MAX_IMAGES") ? atoi(getenv("MDB_CONV_PDL_MAX_IMAGES")) : 16;
  const bool pdl = pdl_on && n_img <= pdl_max_images;
  p.pdl = pdl ? 1 : 0;
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPairThreads);
    cfg.dynamicSmemBytes = op->smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MDB_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<64>, op->maps, p));
  } else {
    conv_pair_kernel<64><<<grid, kPairThreads, op->smem, st>>>(op->maps, p);
  }
  MDB_CUDA(cudaGetLastError());
  if (debug) {
    std::vector<long long> h(static_cast<size_t>(16) * grid);
    MDB_CUDA(cudaStreamSynchronize(st));
    MDB_CUDA(cudaMemcpy(h.data(), d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(d_dbg);
    double p_arr = 0, p_tma = 0, m_commit = 0, m_cacc = 0, m_fence = 0, e_wait = 0, e_proc = 0, e_tot = 0;
    double prod_tot = 0, prod_wait = 0, its = 0, mma_tot = 0, mma_wfull = 0, mma_wacc = 0, tiles = 0, mma_issue = 0;
    int leaders = 0;
    for (int b = 0; b < grid; ++b) {
      p_arr += h[b * 16 + 8]; p_tma += h[b * 16 + 9]; e_wait += h[b * 16 + 13]; e_proc += h[b * 16 + 14]; e_tot += h[b * 16 + 15];
      if (b % 2 == 0) { m_commit += h[b * 16 + 10]; m_cacc += h[b * 16 + 11]; m_fence += h[b * 16 + 12]; }
      prod_tot += h[b * 16 + 0]; prod_wait += h[b * 16 + 1]; its += h[b * 16 + 2];
      if (b % 2 == 0) { mma_tot += h[b * 16 + 4]; mma_wfull += h[b * 16 + 5]; mma_wacc += h[b * 16 + 6]; tiles += h[b * 16 + 7]; mma_issue += h[b * 16 + 3]; ++leaders; }
    }
    fprintf(stderr,
            "[conv dbg] grid %d BN %d stages %d cps %d tiles/cluster %.1f | producer: %.0f cyc total, %.1f%% waiting on empty, "
            "%.0f cyc/stage | mma thread: %.0f cyc total, %.1f%% waiting on full, %.1f%% waiting on acc_empty, %.1f%% in mma issue+commit, %.0f cyc/tile\n",
            grid, p.BN, p.stages, p.cps, tiles / leaders, prod_tot / grid, 100.0 * prod_wait / prod_tot, prod_tot / its,
            mma_tot / leaders, 100.0 * mma_wfull / mma_tot, 100.0 * mma_wacc / mma_tot, 100.0 * mma_issue / mma_tot, mma_tot / tiles);
    fprintf(stderr, "[conv dbg]   per tile (cycles): producer arrive %.0f, tma issue %.0f | mma issue %.0f, commit(empty) %.0f, commit(acc) %.0f, fence %.0f | epilogue warp 0: wait acc_full %.0f, process %.0f, total %.0f\n",
            p_arr / grid / (tiles / leaders), p_tma / grid / (tiles / leaders), mma_issue / tiles, m_commit / tiles, m_cacc / tiles, m_fence / tiles,
            e_wait / grid / (tiles / leaders), e_proc / grid / (tiles / leaders), e_tot / grid / (tiles / leaders));
  }
}

void conv_launch(ConvOp* op, int n_img, cudaStream_t st) { conv_launch_pair(op, n_img, st); }

void conv_free(ConvOp* op) { delete op; }
double conv_macs_per_image(const ConvOp* op) { return op->macs_per_image; }

void pack_weights_khwc(const float* w, int co, int ci, int k, int co_pad, int ci_pad, std::vector<__half>& out) {
  out.assign(static_cast<size_t>(co_pad) * k * k * ci_pad, __float2half(0.f));
  for (int o = 0; o < co; ++o)
    for (int c = 0; c < ci; ++c)
      for (int y = 0; y < k; ++y)
        for (int x = 0; x < k; ++x)
          out[((static_cast<size_t>(o) * k + y) * k + x) * ci_pad + c] =
              __float2half_rn(w[((static_cast<size_t>(o) * ci + c) * k + y) * k + x]);
}

void pack_weights_stem_s2d(const float* w, int co, int co_pad, int ci_pad, std::vector<__half>& out) {
  // 6x6 stride-2 pad-2 conv over RGB == 3x3 stride-1 pad-1 conv over the 2x2 space-to-depth image whose channel is
  // (py*2+px)*3+c.  K1 additionally packs the three column taps into the channel dimension (channel t*16 + ...
  // holds the s2d pixel at column x + t - 1), so the stem runs as a 3 (rows) x 1 conv over 64 channels:
  //   W'[o][a][0][t*16 + (py*2+px)*3 + c] = W[o][c][2a+py][2t+px]
  MDB_REQUIRE(ci_pad == 64, "stem weights are packed for 64 input channels");
  out.assign(static_cast<size_t>(co_pad) * 3 * ci_pad, __float2half(0.f));
  for (int o = 0; o < co; ++o)
    for (int c = 0; c < 3; ++c)
      for (int ky = 0; ky < 6; ++ky)
        for (int kx = 0; kx < 6; ++kx) {
          const int a = ky / 2, py = ky % 2, t = kx / 2, px = kx % 2;
          out[(static_cast<size_t>(o) * 3 + a) * ci_pad + t * 16 + (py * 2 + px) * 3 + c] =
              __float2half_rn(w[((static_cast<size_t>(o) * 3 + c) * 6 + ky) * 6 + kx]);
        }
}

}  // namespace mdb
