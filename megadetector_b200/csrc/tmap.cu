// TMA descriptor construction.  libcuda is not linked (this library must dlopen on a CPU-only box for the
// symbol checks), so cuTensorMapEncodeTiled is resolved through the runtime's driver entry-point query.
#include <cstdlib>
#include <cudaTypedefs.h>

#include "common.h"

namespace mdb {

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  MDB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
  MDB_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// L2 promotion of activation maps (MDB_TMAP_PROMO=128|256|64|0)
static CUtensorMapL2promotion act_promotion() {
  static const char* e = getenv("MDB_TMAP_PROMO");
  const int v = e ? atoi(e) : 128;
  return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                  : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
}

static void encode(CUtensorMap* m, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                   const cuuint32_t* box, int kc, CUtensorMapL2promotion l2) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  MDB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i) MDB_REQUIRE(strides[i] % 16 == 0, "TMA strides must be multiples of 16 bytes");
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), base, dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kc), l2,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u, kc %d)",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1], kc);
    throw CudaError{b};
  }
}

void make_act_tmap(CUtensorMap* m, const View& v, int kc, int tw, int th, int par_y, int par_x) {
  MDB_REQUIRE(!v.fp32, "activation maps are fp16");
  const cuuint64_t px = static_cast<cuuint64_t>(v.Ctot) * 2;  // bytes per pixel of the underlying buffer
  char* base = static_cast<char*>(v.ptr());
  cuuint64_t dims[4], strides[3];
  if (par_y < 0) {
    dims[0] = v.C; dims[1] = v.W; dims[2] = v.H; dims[3] = v.N;
    strides[0] = px; strides[1] = px * v.W; strides[2] = px * v.W * v.H;
  } else {
    MDB_REQUIRE(v.H % 2 == 0 && v.W % 2 == 0, "stride-2 conv needs even input size");
    base += (static_cast<size_t>(par_y) * v.W + par_x) * px;
    dims[0] = v.C; dims[1] = v.W / 2; dims[2] = v.H / 2; dims[3] = v.N;
    strides[0] = px * 2; strides[1] = px * v.W * 2; strides[2] = px * v.W * v.H;
  }
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
  encode(m, base, 4, dims, strides, box, kc, act_promotion());
}

void make_act_tmap_flat(CUtensorMap* m, const View& v, int kc) {
  MDB_REQUIRE(!v.fp32, "activation maps are fp16");
  const cuuint64_t px = static_cast<cuuint64_t>(v.Ctot) * 2;
  const cuuint64_t npix = static_cast<cuuint64_t>(v.N) * v.H * v.W;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(v.C), npix, 1, 1};
  cuuint64_t strides[3] = {px, px * npix, px * npix};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kc), 128, 1, 1};
  encode(m, v.ptr(), 4, dims, strides, box, kc, act_promotion());
}

void make_wgt_tmap(CUtensorMap* m, const void* w, int K, int cout_pad, int kc, int bn) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(cout_pad)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(bn)};
  encode(m, const_cast<void*>(w), 2, dims, strides, box, kc, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

}  // namespace mdb
