// K6 building block shared by the Detect decode kernel (aux.cu) and the stand-alone candidate pass (nms.cu):
// append one 64-bit sort key per candidate to its image's list with warp-aggregated atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mdb {

// key = ~bits(conf) << 32 | anchor: ascending key order == descending confidence, equal confidence: lower anchor first
// (conf > 0, so its bit pattern is monotonic).  EVERY lane of the warp must call this (ok == false: no candidate).
__device__ __forceinline__ void cand_append(bool ok, int img, int row, float conf, unsigned long long* __restrict__ keys,
                                            int32_t* __restrict__ counts, int cap) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t peers = __match_any_sync(0xffffffffu, ok ? img : -1);
  if (ok) {
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (static_cast<int>(lane) == leader) base = atomicAdd(&counts[img], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int pos = base + __popc(peers & ((1u << lane) - 1u));
    if (pos < cap)
      keys[static_cast<size_t>(img) * cap + pos] =
          (static_cast<unsigned long long>(~__float_as_uint(conf)) << 32) | static_cast<uint32_t>(row);
  }
}

}  // namespace mdb
