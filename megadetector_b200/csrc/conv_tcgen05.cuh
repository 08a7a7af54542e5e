// K2: implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces the cuDNN convolutions that the reference triggers through
// `self.model(batch_tensor)` (megadetector/detection/pytorch_detector.py:1313; the module graph is the
// third-party YOLOv5x6 `DetectionModel`, restated in SURVEY.md App. A).  One kernel family covers
// every Conv(+bias)+SiLU of the network:
//
//   GEMM view   D[M = output pixels][N = C_out] = A[M][K = KH*KW*C_in] * W[N][K]^T
//   A operand   never materialised: for each filter tap the TMA engine fetches a
//               {KC channels, tw, th, 1 image} box of the NHWC fp16 activation tensor, shifted by the
//               tap offset; out-of-image elements are zero-filled by the TMA unit (= conv padding).
//               Stride-2 convs use four parity views of the input (one tensor map per (row,col) parity).
//   B operand   fused weights stored [C_out][KH][KW][C_in] fp16 (K-major), 2-D TMA boxes {KC, BN}.
//   D           fp32 accumulators in TMEM, two stages of BN columns so the epilogue of tile i
//               overlaps the MMAs of tile i+1.
//   epilogue    tcgen05.ld -> +bias -> SiLU -> (+residual) -> fp16 (or fp32) NHWC store at a channel
//               offset of the destination buffer (this is how Concat costs nothing).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).  Persistent CTAs walk the tile list.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace mdb {

constexpr int kConvThreads = 192;
constexpr int kConvMaxStages = 12;
constexpr int kTileM = 128;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM column offset between the two accumulator stages

struct ConvParams {
  // output tiling
  int n_img, H_out, W_out;
  int tw, th;  // spatial tile (tw*th <= 128 output pixels = GEMM rows)
  int tiles_x, tiles_y;
  int m_tiles, n_tiles;
  int BN;    // GEMM tile N (multiple of 16, <= 256)
  int Cout;  // real output channels (columns >= Cout are not stored)
  // filter
  int KH, KW, stride, pad;
  int chunks;  // C_in / KC
  int Cin;     // K offset between taps in the weight matrix
  int stages;
  // epilogue
  const float* bias;  // [>= n_tiles*BN]
  void* out;
  long long out_pix_stride;  // elements between consecutive pixels of the destination buffer
  int out_c_off;
  int out_fp32;
  const __half* res;  // optional residual (added after the activation), may be null
  long long res_pix_stride;
  int res_c_off;
  int act;  // 1 = SiLU, 0 = identity
  uint32_t idesc;
};

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

template <int KC>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                  const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  constexpr uint32_t ROW_BYTES = KC * 2;
  constexpr uint32_t A_BYTES = kTileM * ROW_BYTES;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kConvMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kConvMaxStages];
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  // 1024-byte aligned stage ring (required by the 128B swizzle atom)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (static_cast<uint32_t>(p.BN) * ROW_BYTES + 1023u) & ~1023u;
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  const uint32_t tx_bytes = static_cast<uint32_t>(p.tw * p.th) * ROW_BYTES + static_cast<uint32_t>(p.BN) * ROW_BYTES;

  const int S = p.stages;
  const int k_iters = p.KH * p.KW * p.chunks;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (p.stride == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmA3);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full_bar[0], 1);
    mbar_init(&acc_full_bar[1], 1);
    mbar_init(&acc_empty_bar[0], 4);
    mbar_init(&acc_empty_bar[1], 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int txi = m_tile % p.tiles_x;
        const int tyi = (m_tile / p.tiles_x) % p.tiles_y;
        const int img = m_tile / (p.tiles_x * p.tiles_y);
        const int ox0 = txi * p.tw, oy0 = tyi * p.th;
        const int n0 = n_tile * p.BN;
        int ki = 0;
        for (int kh = 0; kh < p.KH; ++kh) {
          for (int kw = 0; kw < p.KW; ++kw) {
            const CUtensorMap* ma = &tmA0;
            int x, y;
            if (p.stride == 1) {
              x = ox0 + kw - p.pad;
              y = oy0 + kh - p.pad;
            } else {
              const int ty_ = kh - p.pad, tx_ = kw - p.pad;
              const int py = ty_ & 1, px = tx_ & 1;
              y = oy0 + (ty_ - py) / 2;
              x = ox0 + (tx_ - px) / 2;
              const int par = py * 2 + px;
              ma = par == 0 ? &tmA0 : (par == 1 ? &tmA1 : (par == 2 ? &tmA2 : &tmA3));
            }
            const int kbase = (kh * p.KW + kw) * p.Cin;
            for (int c = 0; c < p.chunks; ++c, ++ki, ++it) {
              const int s = it % S;
              const uint32_t ph = (it / S) & 1u;
              mbar_wait(&empty_bar[s], ph ^ 1u);
              mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
              uint8_t* sa = smem_raw + (smem_base - smem_u32(smem_raw)) + static_cast<size_t>(s) * stage_bytes;
              tma_load_4d(ma, &full_bar[s], sa, c * KC, x, y, img);
              tma_load_2d(&tmB, &full_bar[s], sa + A_BYTES, kbase + c * KC, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0;
      uint32_t tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u;
        const uint32_t acc_ph = (tcount >> 1) & 1u;
        mbar_wait(&acc_empty_bar[acc], acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int ki = 0; ki < k_iters; ++ki, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + static_cast<uint32_t>(s) * stage_bytes;
          const uint32_t b_addr = a_addr + A_BYTES;
          const uint64_t da = make_kmajor_desc(a_addr, ROW_BYTES);
          const uint64_t db = make_kmajor_desc(b_addr, ROW_BYTES);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            // advance both start addresses by 16 fp16 = 32 bytes (descriptor unit: 16 bytes)
            umma_f16(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), p.idesc,
                     (ki > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs retire
        }
        umma_commit(&acc_full_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const uint32_t quad = warp & 3u;
    const uint32_t row = quad * 32u + lane;
    const int ty = static_cast<int>(row) / p.tw;
    const int tx = static_cast<int>(row) - ty * p.tw;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1u;
      const uint32_t acc_ph = (tcount >> 1) & 1u;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int txi = m_tile % p.tiles_x;
      const int tyi = (m_tile / p.tiles_x) % p.tiles_y;
      const int img = m_tile / (p.tiles_x * p.tiles_y);
      const int ox = txi * p.tw + tx, oy = tyi * p.th + ty;
      const int n0 = n_tile * p.BN;
      const bool valid = (ty < p.th) && (ox < p.W_out) && (oy < p.H_out);
      const long long pix = (static_cast<long long>(img) * p.H_out + oy) * p.W_out + ox;

      mbar_wait(&acc_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + acc * kAccStride;
      for (int j = 0; j < p.BN; j += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(j), v);
        tmem_ld_wait();
        const int col0 = n0 + j;
        if (valid && col0 < p.Cout) {
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float x = __uint_as_float(v[e]) + __ldg(p.bias + col0 + e);
            f[e] = p.act ? silu_f(x) : x;
          }
          if (p.res != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * p.res_pix_stride + p.res_c_off + col0);
            uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
            const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 a = __half22float2(h0[e]);
              float2 b = __half22float2(h1[e]);
              f[2 * e] += a.x;
              f[2 * e + 1] += a.y;
              f[8 + 2 * e] += b.x;
              f[8 + 2 * e + 1] += b.y;
            }
          }
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.out) + pix * p.out_pix_stride + p.out_c_off + col0;
            if (col0 + 16 <= p.Cout) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                reinterpret_cast<float4*>(op)[e] = make_float4(f[4 * e], f[4 * e + 1], f[4 * e + 2], f[4 * e + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (col0 + e < p.Cout) op[e] = f[e];
            }
          } else {
            __half* op = reinterpret_cast<__half*>(p.out) + pix * p.out_pix_stride + p.out_c_off + col0;
            if (col0 + 16 <= p.Cout) {
              uint4 o0, o1;
              __half2* q0 = reinterpret_cast<__half2*>(&o0);
              __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                q0[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                q1[e] = __floats2half2_rn(f[8 + 2 * e], f[8 + 2 * e + 1]);
              }
              reinterpret_cast<uint4*>(op)[0] = o0;
              reinterpret_cast<uint4*>(op)[1] = o1;
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (col0 + e < p.Cout) op[e] = __float2half_rn(f[e]);
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above) -> release the stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace mdb
