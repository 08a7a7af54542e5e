// K2: implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), CTA PAIRS -- tcgen05.mma.cta_group::2, 256 x BN pair tiles.
//
// Replaces the cuDNN convolutions that the reference triggers through `self.model(batch_tensor)`
// (megadetector/detection/pytorch_detector.py:1313; the module graph is the third-party YOLOv5x6 `DetectionModel`,
// restated in SURVEY.md App. A).  One kernel covers every Conv(+bias)+SiLU of the network:
//
//   GEMM view   D[M = output pixels][N = C_out] = A[M][K = KH*KW*C_in] * W[N][K]^T
//   A operand   never materialised: for each filter tap the TMA engine fetches a {64 channels, tw, th, 1 image} box of the
//               NHWC fp16 activation tensor, shifted by the tap offset; out-of-image elements are zero-filled by the TMA
//               unit (= conv padding).  Stride-2 convs use four parity views of the input (one tensor map per parity).
//   B operand   fused weights stored [C_out][KH][KW][C_in] fp16 (K-major), 2-D TMA boxes {64, BN/2} per CTA.
//   D           fp32 accumulators in TMEM (2 stages of up to 256 columns, 4 of up to 128).
//   epilogue    tcgen05.ld -> +bias -> SiLU -> (+residual) -> fp16 (or fp32) NHWC store at a channel offset of the
//               destination buffer (this is how Concat costs nothing).
//
// History (the first version was a single-CTA 128 x BN kernel; profiles/r01_*v1*):
//   * the 128 x BN single-CTA kernel could not keep the tensor pipe fed: 36 KB of operands per 1.3 M MACs and
//     <= 216 KB in flight per SM is not enough at the observed TMA latency (tensor pipe 40 % active on the
//     3x3 320->320 layer, L2 at 37 %, DRAM at 7 %).  A CTA pair shares B: each CTA stages its own 128 A rows
//     and only HALF of the B tile; one elected thread of the leader CTA issues M=256 MMAs that read both
//     CTAs' shared memory.  Operand bytes per MAC drop by 1.4-1.8x, and shared-memory read pressure with them.
//   * every pipeline stage now carries 64 K-elements regardless of the swizzle width (1/2/4 TMA boxes of
//     64/32/16 channels), so the C_in = 80 and C_in = 16 layers no longer pay one mbarrier round trip per MMA.
//   * the epilogue was the bottleneck of every 1x1 layer (17.8 k cycles per 128x160 tile): now 16 warps
//     (four per TMEM lane quadrant, interleaved 16-column chunks), TMEM loads issued one chunk ahead, bias from
//     shared memory, SiLU as ex2.approx + rcp.approx (2 MUFU ops per element).
//
// Warp roles per CTA (576 threads): warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer in the leader CTA),
// warps 2..17 = epilogue.  Barriers: full[s] lives in the LEADER (2 arrivals + bytes of both CTAs),
// empty[s] / acc_full[a] are per CTA and signalled by multicast tcgen05.commit, acc_empty[a] lives in the leader
// (one arrival per epilogue warp of both CTAs).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace mdb {

constexpr int kTileM = 128;     // GEMM rows per CTA (TMEM lanes)
constexpr int kTmemCols = 512;  // the whole TMEM of the SM

// Cycle counters of the role threads (MDB_CONV_DEBUG=1) exist only in -DMDB_CONV_TIMERS builds: even predicated off,
// the extra instructions in the single-thread loops cost 17% on 1x1 layers.
#ifdef MDB_CONV_TIMERS
constexpr bool kTimers = true;
#else
constexpr bool kTimers = false;
#endif
// Timing probes (PairParams::probe) exist only in -DMDB_CONV_PROBES builds
#ifdef MDB_CONV_PROBES
constexpr bool kProbes = true;
#else
constexpr bool kProbes = false;
#endif
#ifndef MDB_EPI_WARPS
#define MDB_EPI_WARPS 16
#endif
constexpr int kPairEpiWarps = MDB_EPI_WARPS;  // a multiple of four: each TMEM lane quadrant gets kPairEpiWarps / 4
#ifndef MDB_EPI_CW
#define MDB_EPI_CW 16
#endif
constexpr int kCW = MDB_EPI_CW;  // accumulator columns an epilogue warp handles per step (16 or 8)
static_assert(kCW == 16 || kCW == 8, "epilogue chunk width must be 16 or 8");
constexpr int kPairThreads = 64 + 32 * kPairEpiWarps;
constexpr int kPairMaxStages = 12;
constexpr int kPairMaxCout = 1280;  // padded output channels per layer (bias staging in shared memory)
constexpr int kHaloRows = 18;                       // th + 2
constexpr int kHaloTw = 8, kHaloTh = 16;
constexpr uint32_t kHaloBox = kHaloRows * kHaloTw * 128;  // bytes of one {64 ch, 8 px, 18 rows} box
constexpr int kHaloMaxA = 4;
#ifndef MDB_HALO_KH_UNROLL
#define MDB_HALO_KH_UNROLL 3  // 3: the nine taps fully unrolled in the MMA issue loop; 1: the filter-row loop stays a loop
#endif
constexpr int kHaloKhUnroll = MDB_HALO_KH_UNROLL;
// single-box halo (stride 1, the default): ONE {64 ch, 10 px, 18 rows} box per chunk serves all nine taps: the kw shift is a
// 128-byte (one pixel) offset of the A descriptor's start address, the kh shift one 1280-byte image row; the eight MMA rows of
// a core-matrix group are the eight pixels of one image row, the stride between groups (SBO) is the 1280-byte row pitch.
// Neither offset is a multiple of the 1024-byte swizzle period: this relies on the 128-byte swizzle being a function of the
// ABSOLUTE shared-memory address (bits 4-6 ^= bits 7-9) on both the TMA write and the MMA read side -- measured: every conv
// parity test passes with the descriptor's base-offset field left 0, and fails when it is set to (start >> 7) & 7
// (profiles/r02_conv_ab_experiments.txt par. 13).  A bytes per chunk and CTA: 23 KB instead of 55 KB.
// Stride 2 works the same way on the four parity views of the input (even/odd rows x even/odd columns): tap (kh, kw) reads view
// (kh != 1, kw != 1) at row offset (kh == 2), column offset (kw == 2), so four boxes of 8|9 px x 16|17 rows (70 KB) replace nine
// 8 x 16 boxes (147 KB).
constexpr uint32_t kHalo1Pitch = (kHaloTw + 2) * 128;
constexpr uint32_t kHalo1Box = kHaloRows * kHalo1Pitch;             // 23040 bytes
constexpr uint32_t kHalo1Stage = (kHalo1Box + 1023u) & ~1023u;      // 23552

struct PairParams {
  int n_img, H_out, W_out;
  int tw, th;
  int tiles_x, tiles_y;
  int m_tiles, pair_m, n_tiles;
  uint32_t mg_ntiles, mg_tx, mg_ty;  // ceil(2^32 / d) for d = n_tiles, tiles_x, tiles_y (0 when d == 1): tile decode without division
  int BN, Cout;
  int taps;          // KH*KW
  int chunks;        // 64-channel chunks per tap
  int total_chunks;  // taps * chunks
  int cps;           // chunks per pipeline stage
  int tail_k;        // 16-channel MMA steps the LAST chunk of a tap needs (1..4): the zero-filled rest of a padded chunk is skipped
  int stages;
  // per filter tap: (dx + 8) | (dy + 8) << 8 | parity_map << 16 -- input offset of the tap relative to the output tile
  // origin (stride 2: in the coordinates of the parity view the tap reads).  One spare entry: the producer reads one ahead.
  uint32_t tap_word[10];
  const float* bias;
  void* out;
  long long out_pix_stride;
  int out_c_off, out_fp32;
  const __half* res;
  long long res_pix_stride;
  int res_c_off;
  int act;
  int wide;        // 1: every epilogue row segment is 32-byte aligned -> 256-bit global accesses
  int fast;        // 1: half output, wide, every 16-column chunk full -> the branch-free epilogue path
  int halo;        // != 0: 3x3 halo main loop (tile 8 x 16 output pixels; A fetched once per chunk and shared by the taps, weights
                   // streamed one filter row per stage); see the tables below
  int pdl;         // launched with programmatic stream serialization: let the next kernel's CTAs start as SMs free up, and wait for
                   // the previous kernel's results only after this CTA's own set-up (barriers, TMEM, bias)
  int a_stages;    // halo mode: depth of the A ring
  // halo mode, table-driven: per chunk `n_boxes` TMA boxes land at box_off[] of the A stage (tensor map box_map[], origin =
  // tile origin + box_dx / box_dy in that map's coordinates); tap (kh, kw) reads its 128 rows at tap_off[kh * 3 + kw] with the
  // 8-row group stride tap_sbo[] (bytes >> 4)
  int n_boxes;
  uint32_t a_stage_bytes, a_tx;
  uint32_t box_off[4];
  int box_map[4], box_dx[4], box_dy[4];
  uint32_t tap_off[9], tap_sbo[9];
  uint32_t idesc;  // M = 256
  long long* dbg;  // per-CTA cycle counters (MDB_CONV_DEBUG=1 in a -DMDB_CONV_TIMERS build), null otherwise
  int backoff_ns;  // > 0: the producer / MMA role threads sleep this long between polls of the barriers they are BLOCKED on when
                   // the epilogue bounds the layer (ring full), instead of spinning: the two spinning role warps were a third of
                   // all issued instructions of a 1x1 layer and share their SM sub-partitions with eight epilogue warps
                   // (profiles/r02_ncu_c1x1_160_*)
  int probe;       // timing probes only (results are garbage), bit mask: 1 = skip TMA loads, 2 = skip MMAs, 4 = epilogue releases
                   // the accumulator without reading it, 8 = no global stores/residual loads, 16 = no SiLU
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-pair rank bit of a shared::cluster address
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per lane and instruction -- half
// the LSU wavefronts of two 128-bit accesses for the pixel-strided epilogue traffic
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  // plain (coherent) load: for single-bottleneck C3 blocks the kernel writes its output in place over the residual
  // (engine.cu Builder::c3), and the non-coherent path must not read memory the same kernel writes
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
// Blackwell's packed FP32 pair instructions (PTX add/mul.f32x2 -> SASS FADD2 / FMUL2): one issue slot for two IEEE-rounded
// results.  The epilogue of the hi-res layers is bound by instruction issue under the power cap (DESIGN.md par. 4), and its
// bias add, scaling, 1 + e, reciprocal sharing and final products all come in natural register pairs (tcgen05.ld and the
// bias LDS.128 deliver adjacent registers).  Results are bit-identical to the scalar form (tools/micro/silu_f32x2_bench.cu:
// same checksums, 9 % less time for the bare math, which is then MUFU-bound).
struct F2 { float x, y; };
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
  unsigned long long pa, pb, pd;
  F2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(pd) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
  return d;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
  unsigned long long pa, pb, pd;
  F2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pd) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
  return d;
}
__device__ __forceinline__ float ex2_approx(float t) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e;
}
// SiLU of four values with ONE reciprocal: 1/a_i from r = 1/(a0 a1 a2 a3) and a few multiplies (a_i = 1 + 2^t_i).  The SFU
// (16 lanes per SM) is the scarcest unit of the epilogue; this cuts it from 2 to 1.25 operations per element.
// t clamped to 30 keeps the product < 2^121; for x < -20.8 the result is |x| * 9.3e-10 instead of something even smaller --
// far below half precision.  Pairs: (x0, x1) and (x2, x3) are adjacent registers.
__device__ __forceinline__ void silu4(float& x0, float& x1, float& x2, float& x3) {
  const F2 k{-1.4426950408889634f, -1.4426950408889634f}, one{1.0f, 1.0f};
  F2 x01{x0, x1}, x23{x2, x3};
  const F2 t01 = mul2(x01, k), t23 = mul2(x23, k);
  const F2 e01 = add2(F2{ex2_approx(fminf(t01.x, 30.0f)), ex2_approx(fminf(t01.y, 30.0f))}, one);
  const F2 e23 = add2(F2{ex2_approx(fminf(t23.x, 30.0f)), ex2_approx(fminf(t23.y, 30.0f))}, one);
  const F2 m = mul2(e01, e23);  // (a0 a2, a1 a3)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m.x * m.y));
  const F2 rr{r * m.y, r * m.x};  // (1 / (a0 a2), 1 / (a1 a3))
  x01 = mul2(x01, mul2(rr, e23));  // x0 / a0, x1 / a1
  x23 = mul2(x23, mul2(rr, e01));  // x2 / a2, x3 / a3
  x0 = x01.x; x1 = x01.y; x2 = x23.x; x3 = x23.y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// TMA descriptors of one convolution: activation maps (dense, or the four stride-2 parity views) and the weight map
struct alignas(64) ConvMaps {
  CUtensorMap a[4];
  CUtensorMap b;
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t x, uint32_t magic) { return magic ? __umulhi(x, magic) : x; }

// pair tile index -> N tile and the (x, y, image) position of THIS CTA's 128-row half
__device__ __forceinline__ void decode_tile(const PairParams& p, int tile, uint32_t rank, int& n_tile, int& txi, int& tyi,
                                            int& img) {
  const uint32_t mt = fastdiv(static_cast<uint32_t>(tile), p.mg_ntiles);
  n_tile = tile - static_cast<int>(mt) * p.n_tiles;
  const uint32_t m_tile = 2u * mt + rank;
  const uint32_t t2 = fastdiv(m_tile, p.mg_tx);
  txi = static_cast<int>(m_tile) - static_cast<int>(t2) * p.tiles_x;
  const uint32_t t3 = fastdiv(t2, p.mg_ty);
  tyi = static_cast<int>(t2) - static_cast<int>(t3) * p.tiles_y;
  img = static_cast<int>(t3);  // >= n_img for the padding tile of an odd tile count: TMA zero-fills, stores are masked
}

__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive (count 1, no tx) on an mbarrier of another CTA of the cluster, given its shared::cluster address.  Default
// (cta-scope) semantics on purpose: a .release.cluster arrive costs a cluster-scope fence per stage in the peer's
// producer and halved the main-loop rate (measured, profiles/r01_conv_cases_v2a_release_cluster_arrive.txt).
__device__ __forceinline__ void mbar_arrive_cluster_addr(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_addr(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma2_load_2d_addr(const CUtensorMap* m, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d_addr(const CUtensorMap* m, uint32_t leader_bar, uint32_t dst, int c0, int c1,
                                                  int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// M=256 MMA with the two shared-memory descriptors given as (low word, common high word): the high word (SBO, version,
// swizzle) never changes, so advancing K or the stage is one 32-bit add
__device__ __forceinline__ void umma2_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with separate high words for A and B (single-box halo: A has its own SBO)
__device__ __forceinline__ void umma2_f16_lo2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc_addr(uint32_t bar_addr) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
      "h"(mask)
      : "memory");
}
// try_wait WITHOUT a suspend-time hint on purpose: with a hint (1 ms, as CUTLASS passes) the hardware parks the thread and
// the wake-up latency cost 5-8% on the 1x1 layers (measured A/B, profiles/r01_conv_ab_experiments.txt par. 4), although the
// polling loops are ~20% of all issued instructions there.
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait on a phase parity; the watchdog counts expired hints and turns a protocol bug into a trap instead of a hung GPU
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait_addr(addr, parity)) return;
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait_addr(addr, parity)) {
    if ((++polls & 63u) == 0 && clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }  // ~2 s of SM clocks
  }
}
// wait that yields the issue slots while blocked: used where being late by `ns` costs nothing (the waiter is ahead of the
// epilogue by a whole accumulator / pipeline ring)
__device__ __forceinline__ void mbar_wait_addr_backoff(uint32_t addr, uint32_t parity, uint32_t ns) {
  if (mbar_try_wait_addr(addr, parity)) return;
  if (ns == 0) { mbar_wait_addr(addr, parity); return; }
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait_addr(addr, parity)) {
    __nanosleep(ns);
    if ((++polls & 63u) == 0 && clock64() - t0 > 4000000000LL) { asm volatile("trap;"); }
  }
}
__device__ __forceinline__ void mbar_wait_addr_timed(uint32_t addr, uint32_t parity, long long& acc) {
  if (!kTimers) { mbar_wait_addr(addr, parity); return; }
  const long long t0 = clock64();
  mbar_wait_addr(addr, parity);
  acc += clock64() - t0;
}

// The three roles run WARP-UNIFORM loops: every lane executes the loop control, the barrier waits and the coordinate
// arithmetic (which ptxas then keeps on the uniform datapath, next to the TMA / MMA operands), and only the instructions
// with side effects are issued by the elected lane.  The first version ran the whole loop inside `if (elect_one())`:
// its ~110 dependent per-thread instructions per stage (R2UR round trips, re-loaded parameters) took 800-1100 cycles
// per stage and bounded every layer with short stages (skeleton probe, profiles/r01_conv_ab_experiments.txt par. 1-2).
template <int KC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ PairParams p) {
  constexpr uint32_t ROW_BYTES = KC * 2;
  constexpr uint32_t A_CHUNK = kTileM * ROW_BYTES;  // 128 rows of one K chunk

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPairMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kPairMaxStages];
  __shared__ __align__(8) uint64_t acc_full_bar[4];
  __shared__ __align__(8) uint64_t acc_empty_bar[4];
  __shared__ __align__(8) uint64_t a_full_bar[kHaloMaxA];
  __shared__ __align__(8) uint64_t a_empty_bar[kHaloMaxA];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_all[kPairMaxCout];  // the layer's whole (padded) bias vector

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t half_n = static_cast<uint32_t>(p.BN) >> 1;
  const uint32_t b_chunk = (half_n * ROW_BYTES + 1023u) & ~1023u;
  const uint32_t chunk_bytes = A_CHUNK + b_chunk;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.cps) * chunk_bytes;  // cps 64-channel chunks per stage
  // bytes that land in ONE CTA per chunk (A box + half-B box); the leader's barrier expects both CTAs' bytes
  const uint32_t tx_per_chunk = static_cast<uint32_t>(p.tw * p.th) * ROW_BYTES + half_n * ROW_BYTES;

  const uint32_t S = static_cast<uint32_t>(p.stages);
  // TMEM accumulator ring: 512 columns hold 2 stages of up to 256 columns or 4 stages of up to 128
  const uint32_t acc_shift = p.BN <= 128 ? 2u : 1u;
  const uint32_t acc_mask = (1u << acc_shift) - 1u;
  const uint32_t acc_stride = p.BN <= 128 ? 128u : 256u;
  const int total_tiles = p.pair_m * p.n_tiles;
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    if (p.tap_word[p.taps - 1] >> 16) {  // stride 2: the parity views
      tma_prefetch_desc(&maps.a[1]);
      tma_prefetch_desc(&maps.a[2]);
      tma_prefetch_desc(&maps.a[3]);
    }
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's instance is the one in use: both producers arrive
      mbar_init(&empty_bar[s], 1);  // multicast commit of the leader's MMA thread
    }
    for (int a = 0; a < kHaloMaxA; ++a) {
      mbar_init(&a_full_bar[a], 2);
      mbar_init(&a_empty_bar[a], 1);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(&acc_full_bar[a], 1);
      mbar_init(&acc_empty_bar[a], 2 * kPairEpiWarps);  // leader's instance: every epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(&tmem_base_smem, kTmemCols);
    tmem_relinquish2();
  }
  for (int i = static_cast<int>(threadIdx.x); i < p.n_tiles * p.BN; i += kPairThreads) bias_all[i] = __ldg(p.bias + i);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (p.pdl) {
    // everything above touched only this layer's weights; activations (TMA loads, residual reads, output stores that may
    // reuse a buffer the previous layer still reads) come after the previous kernel has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  const uint32_t full0 = smem_u32(&full_bar[0]);
  const uint32_t empty0 = smem_u32(&empty_bar[0]);
  const uint32_t a_full0 = smem_u32(&a_full_bar[0]);
  const uint32_t a_empty0 = smem_u32(&a_empty_bar[0]);
  const uint32_t acc_full0 = smem_u32(&acc_full_bar[0]);
  const uint32_t acc_empty0 = smem_u32(&acc_empty_bar[0]);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const bool el = elect_one();
    long long w_empty = 0;
    uint32_t it = 0;
    const long long t_start = kTimers ? clock64() : 0;
    uint32_t s = 0, ph = 0;
    // the LEADER's full barriers: as the mbarrier operand of a cta_group::2 TMA load (peer bit cleared) and as the
    // shared::cluster address the non-leader arrives on
    const uint32_t full0_tma = full0 & kPeerBitMask;
    const uint32_t a_full0_tma = a_full0 & kPeerBitMask;
    const uint32_t full0_remote = mapa_rank(full0, 0);
    const uint32_t a_full0_remote = mapa_rank(a_full0, 0);
    const bool loads = !(kProbes && (p.probe & 1));
    const uint32_t bo = static_cast<uint32_t>(p.backoff_ns);
    if (p.halo) {
      // ---- 3x3 / stride 1: per 64-channel chunk, three {64 ch, 8 px, 18 rows} boxes (one per kw shift) are
      // fetched once and serve all three kh taps (the kh shift is a 1024-byte offset in shared memory);
      // weights stream through their own ring, one filter row per stage.  A bytes per chunk: 55 KB instead of 147 KB.
      const uint32_t a_stage_bytes = p.a_stage_bytes;
      const uint32_t a_tx = p.a_tx;
      const uint32_t b_region = smem_base + static_cast<uint32_t>(p.a_stages) * a_stage_bytes;
      const uint32_t b_bytes = half_n * ROW_BYTES;
      uint32_t as = 0, aph = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        int n_tile, txi, tyi, img;
        decode_tile(p, tile, rank, n_tile, txi, tyi, img);
        const int ox0 = txi * kHaloTw, oy0 = tyi * kHaloTh;
        const int nb0 = n_tile * p.BN + static_cast<int>(rank * half_n);
        for (int c = 0; c < p.chunks; ++c) {
          if (kTimers) mbar_wait_addr_timed(a_empty0 + 8u * as, aph ^ 1u, w_empty);
          else mbar_wait_addr_backoff(a_empty0 + 8u * as, aph ^ 1u, bo);
          if (el) {
            if (leader) mbar_arrive_expect_tx_addr(a_full0 + 8u * as, loads ? 2u * a_tx : 0u);
            else mbar_arrive_cluster_addr(a_full0_remote + 8u * as);
            const uint32_t abase = smem_base + as * a_stage_bytes;
            if (loads) {
              for (int b = 0; b < p.n_boxes; ++b)
                tma2_load_4d_addr(&maps.a[p.box_map[b]], a_full0_tma + 8u * as, abase + p.box_off[b], c * KC, ox0 + p.box_dx[b],
                                  oy0 + p.box_dy[b], img);
            }
          }
          if (++as == static_cast<uint32_t>(p.a_stages)) { as = 0; aph ^= 1u; }
          // weights: one stage per filter row (three taps, 12 MMAs): a stage per tap (4 MMAs, ~160-320 tensor cycles)
          // left the two single-thread roles no slack per barrier round trip
          int kcol = c * KC;
          const int tap_cols = p.chunks * KC;
          for (int kh = 0; kh < 3; ++kh, kcol += 3 * tap_cols) {
            if (kTimers) mbar_wait_addr_timed(empty0 + 8u * s, ph ^ 1u, w_empty);
            else mbar_wait_addr_backoff(empty0 + 8u * s, ph ^ 1u, bo);
            if (el) {
              if (leader) mbar_arrive_expect_tx_addr(full0 + 8u * s, loads ? 6u * b_bytes : 0u);
              else mbar_arrive_cluster_addr(full0_remote + 8u * s);
              if (loads) {
                const uint32_t bdst = b_region + s * 3u * b_chunk;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                  tma2_load_2d_addr(&maps.b, full0_tma + 8u * s, bdst + static_cast<uint32_t>(kw) * b_chunk, kcol + kw * tap_cols, nb0);
              }
            }
            if (++s == S) { s = 0; ph ^= 1u; }
            ++it;
          }
        }
      }
    } else {
      const uint32_t tx_stage = 2u * tx_per_chunk;  // both CTAs' bytes of one chunk
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        int n_tile, txi, tyi, img;
        decode_tile(p, tile, rank, n_tile, txi, tyi, img);
        const int ox0 = txi * p.tw - 8, oy0 = tyi * p.th - 8;  // the tap words carry dx + 8, dy + 8
        const int nb0 = n_tile * p.BN + static_cast<int>(rank * half_n);
        int tap = 0, c = 0, kcol = 0;
        uint32_t w = p.tap_word[0];
        for (int q0 = 0; q0 < p.total_chunks; q0 += p.cps) {
          const int nch = min(p.cps, p.total_chunks - q0);
          if (kTimers) mbar_wait_addr_timed(empty0 + 8u * s, ph ^ 1u, w_empty);
          else mbar_wait_addr_backoff(empty0 + 8u * s, ph ^ 1u, bo);
          if (el) {
            if (leader) mbar_arrive_expect_tx_addr(full0 + 8u * s, loads ? static_cast<uint32_t>(nch) * tx_stage : 0u);
            else mbar_arrive_cluster_addr(full0_remote + 8u * s);
          }
          uint32_t dst = smem_base + s * stage_bytes;
          for (int j = 0; j < nch; ++j, dst += chunk_bytes, kcol += KC) {
            if (el && loads) {
              tma2_load_4d_addr(&maps.a[w >> 16], full0_tma + 8u * s, dst, c * KC, ox0 + static_cast<int>(w & 0xffu),
                                oy0 + static_cast<int>((w >> 8) & 0xffu), img);
              tma2_load_2d_addr(&maps.b, full0_tma + 8u * s, dst + A_CHUNK, kcol, nb0);
            }
            if (++c == p.chunks) {
              c = 0;
              ++tap;
              w = p.tap_word[tap];
            }
          }
          if (++s == S) { s = 0; ph ^= 1u; }
          ++it;
        }
      }
    }
    if (kTimers && p.dbg != nullptr && el) {
      p.dbg[blockIdx.x * 16 + 0] = clock64() - t_start;
      p.dbg[blockIdx.x * 16 + 1] = w_empty;
      p.dbg[blockIdx.x * 16 + 2] = it;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const bool el = elect_one();
      uint32_t tcount = 0;
      long long w_full = 0, w_acc = 0;
      const long long t_start = kTimers ? clock64() : 0;
      uint32_t s = 0, ph = 0;
      uint32_t as = 0, aph = 0;  // halo mode: position in the A ring (continues across tiles)
      const uint64_t desc_tmpl = make_kmajor_desc(0, ROW_BYTES);
      const uint32_t desc_hi = static_cast<uint32_t>(desc_tmpl >> 32);
      const uint32_t desc_lo = static_cast<uint32_t>(desc_tmpl);  // LBO field; the start address is added per use
      const bool mmas = !(kProbes && (p.probe & 2));
      const bool has_tail = p.tail_k != KC / 16;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tcount) {
        const uint32_t acc = tcount & acc_mask;
        const uint32_t acc_ph = (tcount >> acc_shift) & 1u;
        if (kTimers) mbar_wait_addr_timed(acc_empty0 + 8u * acc, acc_ph ^ 1u, w_acc);
        else mbar_wait_addr_backoff(acc_empty0 + 8u * acc, acc_ph ^ 1u, static_cast<uint32_t>(p.backoff_ns));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        // The issuing thread's loop is the network's metronome: one dependent instruction costs it ~10 cycles and an N = 80 MMA
        // lasts 40.  So the nine taps are fully unrolled (their table entries become constant-bank operands), and the
        // "skip the zero-filled K steps" test sits outside the per-MMA code: one uniform branch per chunk, none per MMA
        // (a per-MMA `if (k < nk)` cost 9 % of the whole network: profiles/r02_conv_ab_experiments.txt par. 18).
        if (p.halo) {
          const uint32_t a_stage_bytes = p.a_stage_bytes;
          const uint32_t b_region = smem_base + static_cast<uint32_t>(p.a_stages) * a_stage_bytes;
          const uint32_t a_hi_base = desc_hi & ~0x3FFFu;  // A's high word carries the tap's SBO = its box's image-row pitch
          for (int c = 0; c < p.chunks; ++c) {
            mbar_wait_addr_timed(a_full0 + 8u * as, aph, w_full);
            const uint32_t abase = smem_base + as * a_stage_bytes;
            const bool tail = has_tail && c == p.chunks - 1;
#pragma unroll kHaloKhUnroll
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait_addr_timed(full0 + 8u * s, ph, w_full);
              tc_fence_after();
              if (el) {
                const uint32_t b_stage = b_region + s * 3u * b_chunk;
                if (!tail) {
#pragma unroll
                  for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t a_lo = desc_lo | ((abase + p.tap_off[kh * 3 + kw]) >> 4);
                    const uint32_t a_hi = a_hi_base | p.tap_sbo[kh * 3 + kw];
                    const uint32_t b_lo = desc_lo | ((b_stage + static_cast<uint32_t>(kw) * b_chunk) >> 4);
                    if (mmas) {
#pragma unroll
                      for (int k = 0; k < KC / 16; ++k)
                        umma2_f16_lo2(d_tmem, a_lo + 2u * k, a_hi, b_lo + 2u * k, desc_hi, p.idesc,
                                      (kh | kw | k) != 0 ? 1u : (c != 0 ? 1u : 0u));
                    }
                  }
                } else {
#pragma unroll
                  for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t a_lo = desc_lo | ((abase + p.tap_off[kh * 3 + kw]) >> 4);
                    const uint32_t a_hi = a_hi_base | p.tap_sbo[kh * 3 + kw];
                    const uint32_t b_lo = desc_lo | ((b_stage + static_cast<uint32_t>(kw) * b_chunk) >> 4);
                    if (mmas) {
                      for (int k = 0; k < p.tail_k; ++k)
                        umma2_f16_lo2(d_tmem, a_lo + 2u * k, a_hi, b_lo + 2u * k, desc_hi, p.idesc, (c | kh | kw | k) != 0 ? 1u : 0u);
                    }
                  }
                }
                umma2_commit_mc_addr(empty0 + 8u * s);
              }
              if (++s == S) { s = 0; ph ^= 1u; }
            }
            if (el) umma2_commit_mc_addr(a_empty0 + 8u * as);
            if (++as == static_cast<uint32_t>(p.a_stages)) { as = 0; aph ^= 1u; }
          }
        } else if (!has_tail) {
          for (int q0 = 0; q0 < p.total_chunks; q0 += p.cps) {
            const int nch = min(p.cps, p.total_chunks - q0);
            mbar_wait_addr_timed(full0 + 8u * s, ph, w_full);
            tc_fence_after();
            uint32_t a_lo = desc_lo | ((smem_base + s * stage_bytes) >> 4);
            if (el) {
              for (int j = 0; j < nch; ++j, a_lo += chunk_bytes >> 4) {
                if (mmas) {
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma2_f16_lo(d_tmem, a_lo + 2u * k, a_lo + (A_CHUNK >> 4) + 2u * k, desc_hi, p.idesc,
                                 (q0 | j | k) != 0 ? 1u : 0u);
                }
              }
              umma2_commit_mc_addr(empty0 + 8u * s);
            }
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        } else {
          // C_in not a multiple of 64: the last chunk of every tap issues only tail_k of its four K steps
          int cc = 0;  // chunk index within the tap at the start of the stage
          for (int q0 = 0; q0 < p.total_chunks; q0 += p.cps) {
            const int nch = min(p.cps, p.total_chunks - q0);
            mbar_wait_addr_timed(full0 + 8u * s, ph, w_full);
            tc_fence_after();
            uint32_t a_lo = desc_lo | ((smem_base + s * stage_bytes) >> 4);
            if (el) {
              int c2 = cc;
              for (int j = 0; j < nch; ++j, a_lo += chunk_bytes >> 4) {
                const int nk = c2 == p.chunks - 1 ? p.tail_k : KC / 16;
                if (++c2 == p.chunks) c2 = 0;
                if (mmas) {
                  for (int k = 0; k < nk; ++k)
                    umma2_f16_lo(d_tmem, a_lo + 2u * k, a_lo + (A_CHUNK >> 4) + 2u * k, desc_hi, p.idesc,
                                 (q0 | j | k) != 0 ? 1u : 0u);
                }
              }
              umma2_commit_mc_addr(empty0 + 8u * s);
            }
            cc += nch;  // nch <= 4: a subtract loop, not an integer division, on the issue thread's path
            while (cc >= p.chunks) cc -= p.chunks;
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        }
        if (el) umma2_commit_mc_addr(acc_full0 + 8u * acc);
      }
      if (kTimers && p.dbg != nullptr && el) {
        p.dbg[blockIdx.x * 16 + 4] = clock64() - t_start;
        p.dbg[blockIdx.x * 16 + 5] = w_full;
        p.dbg[blockIdx.x * 16 + 6] = w_acc;
        p.dbg[blockIdx.x * 16 + 7] = tcount;
      }
    }
  } else {
    // ===================== epilogue (warps 2.. of both CTAs) =====================
    const uint32_t ew = warp - 2;          // 0..kPairEpiWarps-1
    const uint32_t quad = warp & 3u;       // TMEM lane quadrant this warp may read
    const uint32_t part = ew >> 2;         // which interleaved set of 16-column chunks
    constexpr int kParts = kPairEpiWarps / 4;
    const uint32_t row = quad * 32u + lane;
    const int ty = static_cast<int>(row) / p.tw;
    const int tx = static_cast<int>(row) - ty * p.tw;
    const int nchunks = p.BN / kCW;
    const uint32_t acc_empty_remote = mapa_rank(acc_empty0, 0);
    uint32_t tcount = 0;
    long long e_wait = 0, e_proc = 0;
    const long long e_start = kTimers ? clock64() : 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tcount) {
      const uint32_t acc = tcount & acc_mask;
      const uint32_t acc_ph = (tcount >> acc_shift) & 1u;
      int n_tile, txi, tyi, img;
      decode_tile(p, tile, rank, n_tile, txi, tyi, img);
      const int ox = txi * p.tw + tx, oy = tyi * p.th + ty;
      const int n0 = n_tile * p.BN;
      const bool valid = (img < p.n_img) && (ty < p.th) && (ox < p.W_out) && (oy < p.H_out);
      const long long pix = (static_cast<long long>(img) * p.H_out + oy) * p.W_out + ox;

      const float* bs = bias_all + n0;

      const long long te0 = kTimers ? clock64() : 0;
      mbar_wait_addr(acc_full0 + 8u * acc, acc_ph);
      tc_fence_after();
      const long long te1 = kTimers ? clock64() : 0;
      const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + acc * acc_stride;

      // ---- fast path: half output, aligned row segments, every chunk full (all but the Detect heads)
      __half* const orow = reinterpret_cast<__half*>(p.out) + pix * p.out_pix_stride + p.out_c_off + n0;
      const __half* const rrow = p.res != nullptr ? p.res + pix * p.res_pix_stride + p.res_c_off + n0 : nullptr;
      // bias + SiLU + residual of one chunk of kCW columns, fp32, in registers
      auto math = [&](const uint32_t (&vv)[kCW], float (&f)[kCW], int jj, const uint4& r0, const uint4& r1, bool have_res) {
        const float4* b4 = reinterpret_cast<const float4*>(bs + jj * kCW);
#pragma unroll
        for (int q = 0; q < kCW / 4; ++q) {
          const float4 bb = b4[q];
          const F2 lo = add2(F2{__uint_as_float(vv[4 * q + 0]), __uint_as_float(vv[4 * q + 1])}, F2{bb.x, bb.y});
          const F2 hi = add2(F2{__uint_as_float(vv[4 * q + 2]), __uint_as_float(vv[4 * q + 3])}, F2{bb.z, bb.w});
          f[4 * q + 0] = lo.x; f[4 * q + 1] = lo.y; f[4 * q + 2] = hi.x; f[4 * q + 3] = hi.y;
        }
        if (p.act && !(kProbes && p.probe & 16)) {
#pragma unroll
          for (int e = 0; e < kCW; e += 4) silu4(f[e], f[e + 1], f[e + 2], f[e + 3]);
        }
        if (have_res) {
          const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
          const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = __half22float2(h0[e]);
            const F2 s0 = add2(F2{f[2 * e], f[2 * e + 1]}, F2{a.x, a.y});
            f[2 * e] = s0.x; f[2 * e + 1] = s0.y;
            if (kCW == 16) {
              const float2 b = __half22float2(h1[e]);
              const F2 s1 = add2(F2{f[(8 + 2 * e) % kCW], f[(8 + 2 * e + 1) % kCW]}, F2{b.x, b.y});
              f[(8 + 2 * e) % kCW] = s1.x; f[(8 + 2 * e + 1) % kCW] = s1.y;
            }
          }
        }
      };
      auto pack = [&](const float (&f)[kCW], uint4& o0, uint4& o1) {
        __half2* q0 = reinterpret_cast<__half2*>(&o0);
        __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          q0[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
          if (kCW == 16) q1[e] = __floats2half2_rn(f[(8 + 2 * e) % kCW], f[(8 + 2 * e + 1) % kCW]);
        }
      };
      auto process_fast = [&](const uint32_t (&vv)[kCW], int jj) {
        if (!valid || (kProbes && p.probe & 4)) return;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
        const bool have_res = rrow != nullptr;
        if (have_res && !(kProbes && (p.probe & 8))) {
          if (kCW == 16) ldg256(rrow + jj * kCW, r0, r1);
          else r0 = *reinterpret_cast<const uint4*>(rrow + jj * kCW);
        }
        float f[kCW];
        math(vv, f, jj, r0, r1, have_res);
        uint4 o0, o1;
        pack(f, o0, o1);
        if (kProbes && p.probe & 8) {
          if (o0.x == 0x7fff7fffu && o0.w == 0x7fff7fffu) *reinterpret_cast<uint4*>(orow + jj * kCW) = o0;  // keeps the math alive, never true
        } else if (kCW == 16) {
          stg256(orow + jj * kCW, o0, o1);
        } else {
          *reinterpret_cast<uint4*>(orow + jj * kCW) = o0;
        }
      };
      // ---- general path: fp32 outputs, partial chunks, unaligned views
      auto process = [&](const uint32_t (&vv)[kCW], int jj) {
        const int col0 = n0 + jj * kCW;
        if (!(valid && col0 < p.Cout) || (kProbes && p.probe & 4)) return;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
        if (p.res != nullptr) {
          const __half* rp = p.res + pix * p.res_pix_stride + p.res_c_off + col0;
          r0 = *reinterpret_cast<const uint4*>(rp);
          if (kCW == 16) r1 = *(reinterpret_cast<const uint4*>(rp) + 1);
        }
        float f[kCW];
        math(vv, f, jj, r0, r1, p.res != nullptr);
        if (p.out_fp32) {
          float* op = reinterpret_cast<float*>(p.out) + pix * p.out_pix_stride + p.out_c_off + col0;
          if (col0 + kCW <= p.Cout) {
#pragma unroll
            for (int e = 0; e < kCW / 4; ++e)
              reinterpret_cast<float4*>(op)[e] = make_float4(f[4 * e], f[4 * e + 1], f[4 * e + 2], f[4 * e + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < kCW; ++e)
              if (col0 + e < p.Cout) op[e] = f[e];
          }
        } else {
          __half* op = reinterpret_cast<__half*>(p.out) + pix * p.out_pix_stride + p.out_c_off + col0;
          if (col0 + kCW <= p.Cout) {
            uint4 o0, o1;
            pack(f, o0, o1);
            reinterpret_cast<uint4*>(op)[0] = o0;
            if (kCW == 16) reinterpret_cast<uint4*>(op)[1] = o1;
          } else {
#pragma unroll
            for (int e = 0; e < kCW; ++e)
              if (col0 + e < p.Cout) op[e] = __float2half_rn(f[e]);
          }
        }
      };
      // TMEM loads run one chunk ahead of the math (two register buffers, statically indexed)
      uint32_t va[kCW], vb[kCW];
      // When the chunk count is not a multiple of the parts (10 chunks at BN = 160, 5 at BN = 80), some warps of a
      // quadrant get one chunk more than the others.  Rotating the assignment from tile to tile by that remainder
      // evens the work out over consecutive tiles; the accumulator ring (2-4 stages) lets the early warps run ahead.
      int j = static_cast<int>((part + tcount * static_cast<uint32_t>(nchunks % kParts)) % static_cast<uint32_t>(kParts));
      if (j < nchunks) tmem_ldN(taddr + static_cast<uint32_t>(j * kCW), va);
      if (p.fast) {
        while (j < nchunks) {
          tmem_ld_wait();
          if (j + kParts < nchunks) tmem_ldN(taddr + static_cast<uint32_t>((j + kParts) * kCW), vb);
          process_fast(va, j);
          j += kParts;
          if (j >= nchunks) break;
          tmem_ld_wait();
          if (j + kParts < nchunks) tmem_ldN(taddr + static_cast<uint32_t>((j + kParts) * kCW), va);
          process_fast(vb, j);
          j += kParts;
        }
      } else {
        while (j < nchunks) {
          tmem_ld_wait();
          if (j + kParts < nchunks) tmem_ldN(taddr + static_cast<uint32_t>((j + kParts) * kCW), vb);
          process(va, j);
          j += kParts;
          if (j >= nchunks) break;
          tmem_ld_wait();
          if (j + kParts < nchunks) tmem_ldN(taddr + static_cast<uint32_t>((j + kParts) * kCW), va);
          process(vb, j);
          j += kParts;
        }
      }
      // all TMEM reads of this warp have completed (tcgen05.wait::ld): hand the accumulator stage back
      tc_fence_before();
      __syncwarp();
      if (kTimers) { e_wait += te1 - te0; e_proc += clock64() - te1; }
      if (lane == 0) {
        if (leader) mbar_arrive(&acc_empty_bar[acc]);
        else mbar_arrive_cluster_addr(acc_empty_remote + 8u * acc);
      }
    }
    if (kTimers && p.dbg != nullptr && ew == 0 && lane == 0) {
      p.dbg[blockIdx.x * 16 + 13] = e_wait;
      p.dbg[blockIdx.x * 16 + 14] = e_proc;
      p.dbg[blockIdx.x * 16 + 15] = clock64() - e_start;
    }
  }

  // both CTAs must be past every remote barrier access before either frees TMEM or exits
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

}  // namespace mdb
