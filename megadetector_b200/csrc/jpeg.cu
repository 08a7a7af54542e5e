// JPEG decode on the GPU (SURVEY.md par. 8f-1): nvJPEG batched decode straight into the buffers K1 reads.
//
// Replaces, for JPEG files, `visualization_utils.load_image` (megadetector/visualization/visualization_utils.py:103-176,
// 306-323: PIL open -> RGB -> EXIF orientation), which `run_detector_batch` calls per image on the host
// (run_detector_batch.py:724) and which caps real-file throughput at ~50-100 images/s per host core.
//
//   * nvJPEG is loaded with dlopen at the first use (libnvjpeg.so.12 ships with the CUDA toolkit): libmdb200.so itself
//     does not depend on it, and a box without it gets a clear error from the JPEG entry points only.
//   * The hardware JPEG engine backend (one batched decode) when it is available and every image of the batch is a baseline
//     JPEG it supports; otherwise the hybrid backend with the batch spread over worker threads (one nvJPEG handle, state and
//     stream each: the Huffman stage runs on the calling CPU thread), output interleaved RGB (NVJPEG_OUTPUT_RGBI; grey
//     images are expanded to RGB like PIL's convert('RGB')).
//   * Four-component (CMYK / YCCK) files are reported as unreadable: the reference rejects PIL mode 'CMYK'
//     (visualization_utils.py:152-155).
//   * EXIF orientation 3 / 6 / 8 is applied by a small rotate kernel (PIL rotate(180 / 270 / 90, expand=True),
//     visualization_utils.py:160-174); mirrored orientations are left untouched, as in the reference (its assert is
//     swallowed by the surrounding try/except).
//
// Pixel parity: nvJPEG and libjpeg-turbo (PIL) use different IDCT / chroma up-sampling arithmetic; decoded pixels differ by
// a few LSB at chroma edges (measured in tests/test_gpu_jpeg.py), which is why this path is compared through the tolerance
// harness and not bit-exactly.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

namespace mdb {

namespace {

struct NvjpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t*, nvjpegPinnedAllocator_t*, unsigned int, nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
  nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char* const*, const size_t*, nvjpegImage_t*, cudaStream_t) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
};

NvjpegApi& api() {
  static NvjpegApi a;
  return a;
}

bool load_api(std::string* err) {
  NvjpegApi& a = api();
  if (a.lib) return true;
  const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (a.lib) break;
  }
  if (!a.lib) {
    *err = "nvJPEG is not available (dlopen libnvjpeg.so.12 failed)";
    return false;
  }
  bool ok = true;
  auto sym = [&](const char* n) {
    void* p = dlsym(a.lib, n);
    if (!p) ok = false;
    return p;
  };
  a.CreateEx = reinterpret_cast<decltype(a.CreateEx)>(sym("nvjpegCreateEx"));
  a.Destroy = reinterpret_cast<decltype(a.Destroy)>(sym("nvjpegDestroy"));
  a.JpegStateCreate = reinterpret_cast<decltype(a.JpegStateCreate)>(sym("nvjpegJpegStateCreate"));
  a.JpegStateDestroy = reinterpret_cast<decltype(a.JpegStateDestroy)>(sym("nvjpegJpegStateDestroy"));
  a.GetImageInfo = reinterpret_cast<decltype(a.GetImageInfo)>(sym("nvjpegGetImageInfo"));
  a.DecodeBatchedInitialize = reinterpret_cast<decltype(a.DecodeBatchedInitialize)>(sym("nvjpegDecodeBatchedInitialize"));
  a.DecodeBatched = reinterpret_cast<decltype(a.DecodeBatched)>(sym("nvjpegDecodeBatched"));
  a.Decode = reinterpret_cast<decltype(a.Decode)>(sym("nvjpegDecode"));
  if (!ok) {
    *err = "libnvjpeg.so lacks an expected symbol";
    dlclose(a.lib);
    a.lib = nullptr;
  }
  return ok;
}

// rotate an interleaved u8 RGB image; rot = 3: 180 degrees, 6: 270 counter-clockwise (= 90 clockwise), 8: 90 counter-clockwise
// (PIL Image.rotate(angle, expand=True) rotates counter-clockwise by `angle`)
__global__ void rotate_rgb_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int rot) {
  const int dw = (rot == 3) ? sw : sh, dh = (rot == 3) ? sh : sw;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(dw) * dh) return;
  const int dx = static_cast<int>(i % dw), dy = static_cast<int>(i / dw);
  int sx, sy;
  if (rot == 3) { sx = sw - 1 - dx; sy = sh - 1 - dy; }
  else if (rot == 6) { sx = dy; sy = sh - 1 - dx; }        // 270 ccw: dst(x, y) = src(y, H-1-x)
  else { sx = sw - 1 - dy; sy = dx; }                       // 90 ccw:  dst(x, y) = src(W-1-y, x)
  const uint8_t* s = src + (static_cast<size_t>(sy) * sw + sx) * 3;
  uint8_t* d = dst + static_cast<size_t>(i) * 3;
  d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

}  // namespace

// EXIF orientation (tag 0x0112 of IFD0 in the APP1 "Exif" segment); 1 when absent or malformed
int jpeg_exif_orientation(const uint8_t* p, size_t n) {
  if (n < 4 || p[0] != 0xFF || p[1] != 0xD8) return 1;
  size_t i = 2;
  while (i + 4 <= n) {
    if (p[i] != 0xFF) return 1;
    const uint8_t m = p[i + 1];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) { i += 2; continue; }
    if (m == 0xDA || m == 0xD9) return 1;  // start of scan / end: no EXIF before the image data
    const size_t len = (static_cast<size_t>(p[i + 2]) << 8) | p[i + 3];
    if (len < 2 || i + 2 + len > n) return 1;
    if (m == 0xE1 && len >= 16 && memcmp(p + i + 4, "Exif\0\0", 6) == 0) {
      const uint8_t* t = p + i + 10;
      const size_t tn = len - 8;
      if (tn < 8) return 1;
      const bool le = t[0] == 'I' && t[1] == 'I';
      if (!le && !(t[0] == 'M' && t[1] == 'M')) return 1;
      auto u16 = [&](size_t o) -> uint32_t { return le ? (t[o] | (t[o + 1] << 8)) : ((t[o] << 8) | t[o + 1]); };
      auto u32 = [&](size_t o) -> uint32_t {
        return le ? (t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | (static_cast<uint32_t>(t[o + 3]) << 24))
                  : ((static_cast<uint32_t>(t[o]) << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]);
      };
      const size_t ifd = u32(4);
      if (ifd + 2 > tn) return 1;
      const uint32_t cnt = u16(ifd);
      for (uint32_t k = 0; k < cnt; ++k) {
        const size_t e = ifd + 2 + 12 * static_cast<size_t>(k);
        if (e + 12 > tn) return 1;
        if (u16(e) == 0x0112) {
          const uint32_t v = u16(e + 8);
          return (v >= 1 && v <= 8) ? static_cast<int>(v) : 1;
        }
      }
      return 1;
    }
    i += 2 + len;
  }
  return 1;
}

// The hybrid backend does the Huffman stage on the calling CPU thread (and nvjpegDecodeBatched only moves it to the GPU for
// batches > 100 images), so one thread decodes ~450 1080p images/s.  The batch is therefore spread over worker threads, each
// with its own nvJPEG handle, state and stream (nvjpegDecode is thread-safe across states).
struct JpegWorker {
  nvjpegHandle_t h = nullptr;
  nvjpegJpegState_t s = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
};

struct JpegDecoder {
  nvjpegHandle_t h_hw = nullptr, h_hybrid = nullptr;
  nvjpegJpegState_t s_hw = nullptr, s_hybrid = nullptr;
  int init_hw = 0, init_hybrid = 0;  // batch size the state was initialised for
  uint8_t* d_buf = nullptr;          // decoded images (and rotated copies), packed
  size_t d_cap = 0;
  bool hw_tried = false;
  std::vector<JpegWorker> workers;
};

JpegDecoder* jpeg_create() { return new JpegDecoder(); }

void jpeg_destroy(JpegDecoder* d) {
  if (!d) return;
  NvjpegApi& a = api();
  for (JpegWorker& w : d->workers) {
    if (a.lib && w.s) a.JpegStateDestroy(w.s);
    if (a.lib && w.h) a.Destroy(w.h);
    if (w.ev) cudaEventDestroy(w.ev);
    if (w.st) cudaStreamDestroy(w.st);
  }
  if (a.lib) {
    if (d->s_hw) a.JpegStateDestroy(d->s_hw);
    if (d->s_hybrid) a.JpegStateDestroy(d->s_hybrid);
    if (d->h_hw) a.Destroy(d->h_hw);
    if (d->h_hybrid) a.Destroy(d->h_hybrid);
  }
  if (d->d_buf) cudaFree(d->d_buf);
  delete d;
}

// Decodes n JPEG streams on `st`.  out[i] receives the device pointer / size of the decoded, orientation-corrected RGB image
// (dense rows), or ptr == nullptr when the stream cannot be decoded.  The buffers belong to the decoder and stay valid until
// the next call.  Returns which backend decoded the batch: 1 = hardware engine, 2 = GPU hybrid.
int jpeg_decode_batch(JpegDecoder* d, const uint8_t* const* jpeg, const size_t* nbytes, int n, JpegImage* out, cudaStream_t st) {
  std::string err;
  if (!load_api(&err)) throw CudaError{err};
  NvjpegApi& a = api();
  if (!d->h_hybrid) {
    if (a.CreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &d->h_hybrid) != NVJPEG_STATUS_SUCCESS)
      throw CudaError{"nvjpegCreateEx(GPU_HYBRID) failed"};
    if (a.JpegStateCreate(d->h_hybrid, &d->s_hybrid) != NVJPEG_STATUS_SUCCESS) throw CudaError{"nvjpegJpegStateCreate failed"};
  }
  if (!d->hw_tried) {
    d->hw_tried = true;
    if (a.CreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &d->h_hw) == NVJPEG_STATUS_SUCCESS) {
      if (a.JpegStateCreate(d->h_hw, &d->s_hw) != NVJPEG_STATUS_SUCCESS) { a.Destroy(d->h_hw); d->h_hw = nullptr; }
    } else {
      d->h_hw = nullptr;
    }
  }
  // sizes, layout
  std::vector<int> w(n, 0), h(n, 0), rot(n, 1);
  std::vector<size_t> off(n, 0), off_rot(n, 0);
  std::vector<int> good;
  size_t total = 0;
  bool baseline_like = true;
  for (int i = 0; i < n; ++i) {
    out[i] = JpegImage{nullptr, 0, 0};
    if (!jpeg[i] || nbytes[i] < 4) continue;
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    if (a.GetImageInfo(d->h_hybrid, jpeg[i], nbytes[i], &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) continue;
    if (nc != 1 && nc != 3) continue;  // CMYK / YCCK: PIL mode 'CMYK' is rejected by the reference
    if (ws[0] <= 0 || hs[0] <= 0 || ws[0] > 32768 || hs[0] > 32768) continue;
    if (ss == NVJPEG_CSS_410 || ss == NVJPEG_CSS_411 || ss == NVJPEG_CSS_UNKNOWN) baseline_like = false;
    w[i] = ws[0]; h[i] = hs[0];
    const int o = jpeg_exif_orientation(jpeg[i], nbytes[i]);
    rot[i] = (o == 3 || o == 6 || o == 8) ? o : 1;
    const size_t bytes = ((static_cast<size_t>(w[i]) * h[i] * 3 + 255) / 256) * 256;
    off[i] = total;
    total += bytes;
    if (rot[i] != 1) { off_rot[i] = total; total += bytes; }
    good.push_back(i);
  }
  if (good.empty()) return 0;
  if (total > d->d_cap) {
    if (d->d_buf) MDB_CUDA(cudaFree(d->d_buf));
    d->d_buf = nullptr; d->d_cap = 0;
    const size_t want = total + total / 4 + 4096;
    MDB_CUDA(cudaMalloc(&d->d_buf, want));
    d->d_cap = want;
  }
  const int m = static_cast<int>(good.size());
  std::vector<const unsigned char*> ptrs(m);
  std::vector<size_t> lens(m);
  std::vector<nvjpegImage_t> dst(m);
  for (int k = 0; k < m; ++k) {
    const int i = good[k];
    ptrs[k] = jpeg[i];
    lens[k] = nbytes[i];
    memset(&dst[k], 0, sizeof(nvjpegImage_t));
    dst[k].channel[0] = d->d_buf + off[i];
    dst[k].pitch[0] = static_cast<size_t>(w[i]) * 3;
  }
  int backend = 0;
  auto run = [&](nvjpegHandle_t hnd, nvjpegJpegState_t stt, int* init) -> bool {
    if (*init != m) {
      if (a.DecodeBatchedInitialize(hnd, stt, m, 1, NVJPEG_OUTPUT_RGBI) != NVJPEG_STATUS_SUCCESS) return false;
      *init = m;
    }
    return a.DecodeBatched(hnd, stt, ptrs.data(), lens.data(), dst.data(), st) == NVJPEG_STATUS_SUCCESS;
  };
  static const bool no_hw = getenv("MDB_JPEG_NO_HW") && getenv("MDB_JPEG_NO_HW")[0] == '1';
  if (d->h_hw && baseline_like && !no_hw && run(d->h_hw, d->s_hw, &d->init_hw)) backend = 1;
  if (backend == 0) {
    // hybrid backend, one image per nvjpegDecode call, spread over worker threads
    d->init_hw = 0;
    backend = 2;
    static const int n_threads = [] {
      const char* e = getenv("MDB_JPEG_THREADS");
      int t = e ? atoi(e) : 16;
      return t < 1 ? 1 : (t > 32 ? 32 : t);
    }();
    const int T = std::min(n_threads, m);
    int device = 0;
    MDB_CUDA(cudaGetDevice(&device));
    while (static_cast<int>(d->workers.size()) < T) {
      JpegWorker wk;
      if (a.CreateEx(NVJPEG_BACKEND_HYBRID, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &wk.h) != NVJPEG_STATUS_SUCCESS)
        throw CudaError{"nvjpegCreateEx(HYBRID) failed"};
      if (a.JpegStateCreate(wk.h, &wk.s) != NVJPEG_STATUS_SUCCESS) throw CudaError{"nvjpegJpegStateCreate failed"};
      MDB_CUDA(cudaStreamCreateWithFlags(&wk.st, cudaStreamNonBlocking));
      MDB_CUDA(cudaEventCreateWithFlags(&wk.ev, cudaEventDisableTiming));
      d->workers.push_back(wk);
    }
    std::vector<char> failed(m, 0);
    auto work = [&](int t) {
      cudaSetDevice(device);
      JpegWorker& wk = d->workers[t];
      for (int k = t; k < m; k += T)
        if (a.Decode(wk.h, wk.s, ptrs[k], lens[k], NVJPEG_OUTPUT_RGBI, &dst[k], wk.st) != NVJPEG_STATUS_SUCCESS) failed[k] = 1;
      cudaEventRecord(wk.ev, wk.st);
    };
    if (T == 1) {
      work(0);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
    for (int t = 0; t < T; ++t) MDB_CUDA(cudaStreamWaitEvent(st, d->workers[t].ev, 0));
    for (int k = 0; k < m; ++k)
      if (failed[k]) w[good[k]] = 0;  // e.g. a truncated stream: reported as an unreadable image
  }
  for (int i : good) {
    if (w[i] == 0) continue;
    if (rot[i] == 1) {
      out[i] = JpegImage{d->d_buf + off[i], h[i], w[i]};
    } else {
      const long long px = static_cast<long long>(w[i]) * h[i];
      rotate_rgb_kernel<<<static_cast<unsigned>((px + 255) / 256), 256, 0, st>>>(d->d_buf + off[i], w[i], h[i],
                                                                                 d->d_buf + off_rot[i], rot[i]);
      MDB_CUDA(cudaGetLastError());
      out[i] = (rot[i] == 3) ? JpegImage{d->d_buf + off_rot[i], h[i], w[i]} : JpegImage{d->d_buf + off_rot[i], w[i], h[i]};
    }
  }
  return backend;
}

}  // namespace mdb
