"""
ctypes binding of include/mdb200.h (libmdb200.so).  This is the only place Python touches the C ABI.

There is no fallback: if the shared library is missing or cannot be loaded, importing an engine raises.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get('MDB_LIB_VARIANT', '')  # experiment builds only (see build.py)
LIB_PATH = os.path.join(_HERE, 'libmdb200{}.so'.format('_' + _VARIANT if _VARIANT else ''))

MDB_OK = 0
MDB_IMG_OK, MDB_IMG_BAD_INPUT, MDB_IMG_INFER_FAILED = 0, 1, 2
MDB_FETCH_STEM_INPUT, MDB_FETCH_PRED, MDB_FETCH_LAYER = 1, 2, 3

EXPORTS = [
    'mdb_create', 'mdb_load_weights', 'mdb_destroy', 'mdb_last_error', 'mdb_infer_batch', 'mdb_batch_begin', 'mdb_batch_next', 'mdb_infer_preprocessed',
    'mdb_letterbox_geometry', 'mdb_set_image_size', 'mdb_debug_fetch', 'mdb_last_timing', 'mdb_bench_resident', 'mdb_profile_steps', 'mdb_set_pipeline_split', 'mdb_launch_count', 'mdb_op_letterbox',
    'mdb_op_conv2d', 'mdb_op_nms', 'mdb_alloc_pinned', 'mdb_free_pinned', 'mdb_version',
    'mdb_infer_jpeg_batch', 'mdb_op_decode_jpeg', 'mdb_infer_preprocessed_rp', 'mdb_set_augment',
    'mdb_debug_halo_layout', 'mdb_debug_set_trace', 'mdb_debug_trace_names', 'mdb_debug_fetch_trace', 'mdb_op_upsample2x', 'mdb_op_sppf_pool', 'mdb_op_decode',
]


class MdbArch(C.Structure):
    _fields_ = [('ch', C.c_int32 * 6), ('n_backbone', C.c_int32 * 5), ('n_head', C.c_int32), ('nc', C.c_int32),
                ('anchors_px', (C.c_float * 2) * 3 * 4)]


class MdbConfig(C.Structure):
    _fields_ = [('struct_size', C.c_int32), ('device', C.c_int32), ('max_batch', C.c_int32), ('img_size', C.c_int32),
                ('stride', C.c_int32), ('letterbox_auto', C.c_int32), ('letterbox_scaleup', C.c_int32),
                ('iou_thres', C.c_float), ('max_det', C.c_int32), ('compat_mode', C.c_int32),
                ('use_ceil_for_resize', C.c_int32), ('arch', MdbArch)]


class MdbTensorDesc(C.Structure):
    _fields_ = [('name', C.c_char_p), ('data', C.POINTER(C.c_float)), ('ndim', C.c_int32), ('shape', C.c_int64 * 4)]


_lib = None


def load_library():
    """dlopen libmdb200.so (building it first if nvcc is here and the sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not (_VARIANT and os.path.exists(LIB_PATH)):  # an experiment variant is frozen once built (A/B against the tree)
        from . import build as _b
        try:  # rebuild when the sources changed and nvcc is here (a no-op stamp check otherwise)
            _b.build()
        except Exception as e:
            # Never load a library that is older than the sources it is bound against: without nvcc a prebuilt .so
            # whose stamp matches is fine, anything else is an error.
            stamp_file = os.path.join(_b.OBJ_DIR, 'stamp')
            ok = os.path.exists(LIB_PATH) and os.path.exists(stamp_file) and open(stamp_file).read() == _b._stamp()
            if not ok:
                raise RuntimeError('libmdb200.so is missing or stale and could not be rebuilt: {}'.format(e))
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libmdb200.so is missing ({}): build it with `python -m megadetector_b200.build`. '
            'megadetector_b200 has no CPU / PyTorch fallback.'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp = C.c_void_p
    lib.mdb_create.argtypes = [P(vp), P(MdbConfig)]
    lib.mdb_create.restype = C.c_int
    lib.mdb_load_weights.argtypes = [vp, P(MdbTensorDesc), C.c_int32]
    lib.mdb_load_weights.restype = C.c_int
    lib.mdb_destroy.argtypes = [vp]
    lib.mdb_destroy.restype = None
    lib.mdb_last_error.argtypes = [vp]
    lib.mdb_last_error.restype = C.c_char_p
    lib.mdb_infer_batch.argtypes = [vp, P(vp), P(C.c_int32), P(C.c_int32), P(C.c_int64), C.c_int32, C.c_float,
                                    vp, vp, vp, vp]
    lib.mdb_infer_batch.restype = C.c_int
    lib.mdb_batch_begin.argtypes = lib.mdb_infer_batch.argtypes
    lib.mdb_batch_begin.restype = C.c_int
    lib.mdb_batch_next.argtypes = [vp, P(C.c_int32), C.c_int32, P(C.c_int32)]
    lib.mdb_batch_next.restype = C.c_int
    lib.mdb_infer_preprocessed.argtypes = [vp, P(vp), P(C.c_int32), P(C.c_int32), P(C.c_int32), P(C.c_int32),
                                           C.c_int32, C.c_float, vp, vp, vp, vp]
    lib.mdb_infer_preprocessed.restype = C.c_int
    lib.mdb_letterbox_geometry.argtypes = [P(MdbConfig), C.c_int32, C.c_int32, P(C.c_int32), P(C.c_double)]
    lib.mdb_letterbox_geometry.restype = C.c_int
    lib.mdb_set_image_size.argtypes = [vp, C.c_int32]
    lib.mdb_set_image_size.restype = C.c_int
    lib.mdb_debug_fetch.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_size_t, P(C.c_int64)]
    lib.mdb_debug_fetch.restype = C.c_int
    lib.mdb_last_timing.argtypes = [vp, P(C.c_float)]
    lib.mdb_last_timing.restype = C.c_int
    lib.mdb_bench_resident.argtypes = [vp, C.c_int32, C.c_float, C.c_int32, P(C.c_float)]
    lib.mdb_bench_resident.restype = C.c_int
    lib.mdb_profile_steps.argtypes = [vp, C.c_int32, vp, vp, vp, C.c_int32, C.c_int32, P(C.c_int32)]
    lib.mdb_profile_steps.restype = C.c_int
    lib.mdb_set_pipeline_split.argtypes = [vp, C.c_int32]
    lib.mdb_set_pipeline_split.restype = C.c_int
    lib.mdb_launch_count.argtypes = [vp]
    lib.mdb_launch_count.restype = C.c_int64
    lib.mdb_op_letterbox.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int64, vp, P(C.c_int32)]
    lib.mdb_op_letterbox.restype = C.c_int
    lib.mdb_op_conv2d.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, vp, C.c_int32, vp, C.c_int32, P(C.c_float)]
    lib.mdb_op_conv2d.restype = C.c_int
    lib.mdb_op_nms.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_double, C.c_int32, vp, vp, vp]
    lib.mdb_op_nms.restype = C.c_int
    lib.mdb_set_augment.argtypes = [vp, C.c_int32]
    lib.mdb_set_augment.restype = C.c_int
    lib.mdb_infer_preprocessed_rp.argtypes = [vp, P(vp), P(C.c_int32), P(C.c_int32), P(C.c_int32), P(C.c_int32), vp,
                                              C.c_int32, C.c_float, vp, vp, vp, vp]
    lib.mdb_infer_preprocessed_rp.restype = C.c_int
    lib.mdb_infer_jpeg_batch.argtypes = [vp, P(vp), P(C.c_size_t), C.c_int32, C.c_float, vp, vp, vp, vp, vp]
    lib.mdb_infer_jpeg_batch.restype = C.c_int
    lib.mdb_op_decode_jpeg.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, P(C.c_int32), P(C.c_int32)]
    lib.mdb_op_decode_jpeg.restype = C.c_int
    lib.mdb_debug_set_trace.argtypes = [vp, C.c_char_p]
    lib.mdb_debug_set_trace.restype = C.c_int
    lib.mdb_debug_halo_layout.argtypes = [C.c_int32, P(C.c_int32)]
    lib.mdb_debug_halo_layout.restype = C.c_int
    lib.mdb_debug_trace_names.argtypes = [vp, vp, C.c_size_t, P(C.c_int32)]
    lib.mdb_debug_trace_names.restype = C.c_int
    lib.mdb_debug_fetch_trace.argtypes = [vp, C.c_char_p, C.c_int32, vp, C.c_size_t, P(C.c_int64)]
    lib.mdb_debug_fetch_trace.restype = C.c_int
    lib.mdb_op_upsample2x.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.mdb_op_upsample2x.restype = C.c_int
    lib.mdb_op_sppf_pool.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.mdb_op_sppf_pool.restype = C.c_int
    lib.mdb_op_decode.argtypes = [vp, P(vp), P(C.c_int32), P(C.c_int32), C.c_int32, vp]
    lib.mdb_op_decode.restype = C.c_int
    lib.mdb_alloc_pinned.argtypes = [C.c_size_t]
    lib.mdb_alloc_pinned.restype = vp
    lib.mdb_free_pinned.argtypes = [vp]
    lib.mdb_free_pinned.restype = None
    lib.mdb_version.argtypes = []
    lib.mdb_version.restype = C.c_char_p
    _lib = lib
    return lib


DEFAULT_ANCHORS_PX = [
    [19, 27, 44, 40, 38, 94],
    [96, 68, 86, 152, 180, 137],
    [140, 301, 303, 264, 238, 542],
    [436, 615, 739, 380, 925, 792],
]


def make_config(device=0, max_batch=8, img_size=1280, stride=64, letterbox_auto=True, letterbox_scaleup=True,
                iou_thres=0.45, max_det=300, ch=(80, 160, 320, 640, 960, 1280), n_backbone=(4, 8, 12, 4, 4), n_head=4,
                nc=3, anchors_px=None, compat_mode=0, use_ceil_for_resize=False):
    cfg = MdbConfig()
    cfg.struct_size = C.sizeof(MdbConfig)
    cfg.device = device
    cfg.max_batch = max_batch
    cfg.img_size = img_size
    cfg.stride = stride
    cfg.letterbox_auto = int(bool(letterbox_auto))
    cfg.letterbox_scaleup = int(bool(letterbox_scaleup))
    cfg.iou_thres = iou_thres
    cfg.max_det = max_det
    cfg.compat_mode = int(compat_mode)
    cfg.use_ceil_for_resize = int(bool(use_ceil_for_resize))
    for i in range(6):
        cfg.arch.ch[i] = int(ch[i])
    for i in range(5):
        cfg.arch.n_backbone[i] = int(n_backbone[i])
    cfg.arch.n_head = int(n_head)
    cfg.arch.nc = int(nc)
    anchors_px = anchors_px or DEFAULT_ANCHORS_PX
    for l in range(4):
        for a in range(3):
            cfg.arch.anchors_px[l][a][0] = float(anchors_px[l][2 * a])
            cfg.arch.anchors_px[l][a][1] = float(anchors_px[l][2 * a + 1])
    return cfg


def letterbox_geometry(cfg, h, w):
    """Pure host arithmetic (works without a GPU): dict mirroring yolov5 letterbox() geometry."""
    lib = load_library()
    out = (C.c_int32 * 8)()
    rp = (C.c_double * 3)()
    rc = lib.mdb_letterbox_geometry(C.byref(cfg), h, w, out, rp)
    if rc != MDB_OK:
        raise ValueError('mdb_letterbox_geometry failed ({})'.format(rc))
    return {'new_unpad': (out[0], out[1]), 'top': out[2], 'bottom': out[3], 'left': out[4], 'right': out[5],
            'out_hw': (out[6], out[7]), 'ratio': (rp[0], rp[0]), 'pad': (rp[1], rp[2])}


def halo_layout(mode):
    """Addressing tables of K2's 3x3 halo main loop (pure host arithmetic; mdb_debug_halo_layout): dict of lists."""
    lib = load_library()
    o = (C.c_int32 * 64)()
    rc = lib.mdb_debug_halo_layout(mode, o)
    if rc != MDB_OK:
        raise ValueError('mdb_debug_halo_layout({}) failed ({})'.format(mode, rc))
    n = o[0]
    return {'n_boxes': n, 'a_tx': o[1], 'a_stage_bytes': o[2], 'box_off': list(o[4:4 + n]), 'box_map': list(o[8:8 + n]),
            'box_par': list(o[12:12 + n]), 'box_dx': list(o[16:16 + n]), 'box_dy': list(o[20:20 + n]),
            'box_w': list(o[24:24 + n]), 'box_h': list(o[28:28 + n]), 'tap_off': list(o[32:41]),
            'tap_sbo': [v << 4 for v in o[41:50]]}


class PinnedArray:
    """numpy view over page-locked host memory owned by libmdb200."""

    def __init__(self, shape, dtype=np.uint8):
        lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = lib.mdb_alloc_pinned(self.nbytes)
        if not self._ptr:
            raise MemoryError('mdb_alloc_pinned({}) failed'.format(self.nbytes))
        buf = (C.c_uint8 * self.nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, '_ptr', None):
                load_library().mdb_free_pinned(self._ptr)
                self._ptr = None
        except Exception:
            pass


class Engine:
    """One mdb_engine handle (one GPU)."""

    def __init__(self, cfg):
        self.lib = load_library()
        self.cfg = cfg
        self.max_det = cfg.max_det
        self.no = cfg.arch.nc + 5
        h = C.c_void_p()
        rc = self.lib.mdb_create(C.byref(h), C.byref(cfg))
        if rc != MDB_OK:
            raise RuntimeError('mdb_create failed ({}): {}'.format(rc, (self.lib.mdb_last_error(None) or b'').decode()))
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, 'h', None):
            self.lib.mdb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != MDB_OK:
            raise RuntimeError('{} failed ({}): {}'.format(what, rc, (self.lib.mdb_last_error(self.h) or b'').decode()))

    def last_error(self):
        return (self.lib.mdb_last_error(self.h) or b'').decode()

    def load_weights(self, weights):
        """weights: dict name -> float32 array-like (numpy or torch CPU tensor), fused-yolov5 naming."""
        descs = (MdbTensorDesc * len(weights))()
        keep = []
        for i, (name, t) in enumerate(weights.items()):
            a = np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, 'detach') else t, dtype=np.float32)
            keep.append(a)
            descs[i].name = name.encode()
            descs[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            descs[i].ndim = a.ndim
            for d in range(a.ndim):
                descs[i].shape[d] = a.shape[d]
        self._check(self.lib.mdb_load_weights(self.h, descs, len(weights)), 'mdb_load_weights')

    def infer_batch(self, images, conf_thres, want_anchors=True):
        """
        images: list of HxWx3 uint8 arrays (C-contiguous rows; row stride taken from .strides).
        Returns (det [n,max_det,6] float32, count [n] int32, anchors [n,max_det] int32 or None, status [n] int32).
        """
        n = len(images)
        ptrs = (C.c_void_p * n)()
        hs = (C.c_int32 * n)()
        ws = (C.c_int32 * n)()
        st = (C.c_int64 * n)()
        for i, im in enumerate(images):
            ok = isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3 and \
                im.strides[2] == 1 and im.strides[1] == 3
            if not ok:
                ptrs[i] = None
                hs[i] = ws[i] = 0
                continue
            ptrs[i] = im.ctypes.data
            hs[i], ws[i] = im.shape[0], im.shape[1]
            st[i] = im.strides[0]
        det = np.zeros((n, self.max_det, 6), np.float32)
        cnt = np.zeros((n,), np.int32)
        anc = np.full((n, self.max_det), -1, np.int32) if want_anchors else None
        status = np.zeros((n,), np.int32)
        rc = self.lib.mdb_infer_batch(self.h, ptrs, hs, ws, st, n, float(conf_thres), det.ctypes.data, cnt.ctypes.data,
                                      anc.ctypes.data if anc is not None else None, status.ctypes.data)
        self._check(rc, 'mdb_infer_batch')
        return det, cnt, anc, status

    def infer_batch_stream(self, images, conf_thres):
        """
        Generator form of infer_batch: yields (indices, det, count, status) each time a chunk of the batch is final
        (det/count/status are the full-batch arrays, valid at `indices`), so the caller can post-process finished
        images while the GPU works on the rest.
        """
        n = len(images)
        ptrs = (C.c_void_p * n)()
        hs = (C.c_int32 * n)()
        ws = (C.c_int32 * n)()
        st = (C.c_int64 * n)()
        for i, im in enumerate(images):
            ok = isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3 and \
                im.strides[2] == 1 and im.strides[1] == 3
            if not ok:
                ptrs[i] = None
                hs[i] = ws[i] = 0
                continue
            ptrs[i] = im.ctypes.data
            hs[i], ws[i] = im.shape[0], im.shape[1]
            st[i] = im.strides[0]
        det = np.zeros((n, self.max_det, 6), np.float32)
        cnt = np.zeros((n,), np.int32)
        status = np.zeros((n,), np.int32)
        self._check(self.lib.mdb_batch_begin(self.h, ptrs, hs, ws, st, n, float(conf_thres), det.ctypes.data,
                                             cnt.ctypes.data, None, status.ctypes.data), 'mdb_batch_begin')
        done = (C.c_int32 * max(1, n))()
        nd = C.c_int32(0)
        finished = False
        try:
            while True:
                self._check(self.lib.mdb_batch_next(self.h, done, max(1, n), C.byref(nd)), 'mdb_batch_next')
                if nd.value == 0:
                    finished = True
                    break
                yield [done[i] for i in range(nd.value)], det, cnt, status
        finally:
            # never leave a half-consumed batch inside the engine (its output arrays die with this frame)
            while not finished:
                if self.lib.mdb_batch_next(self.h, done, max(1, n), C.byref(nd)) != MDB_OK or nd.value == 0:
                    finished = True

    def infer_preprocessed(self, images, src_shapes, conf_thres, want_anchors=True, gain_pad=None):
        """gain_pad: optional [n][3] (gain, pad_x, pad_y) = the ratio_pad of scale_coords ("modern" mode)."""
        n = len(images)
        ptrs = (C.c_void_p * n)()
        lh = (C.c_int32 * n)()
        lw = (C.c_int32 * n)()
        sh = (C.c_int32 * n)()
        sw = (C.c_int32 * n)()
        keep = []
        for i, im in enumerate(images):
            a = np.ascontiguousarray(im, dtype=np.uint8)
            keep.append(a)
            ptrs[i] = a.ctypes.data
            lh[i], lw[i] = a.shape[0], a.shape[1]
            sh[i], sw[i] = int(src_shapes[i][0]), int(src_shapes[i][1])
        det = np.zeros((n, self.max_det, 6), np.float32)
        cnt = np.zeros((n,), np.int32)
        anc = np.full((n, self.max_det), -1, np.int32) if want_anchors else None
        status = np.zeros((n,), np.int32)
        gp = None if gain_pad is None else np.ascontiguousarray(gain_pad, np.float32).reshape(n, 3)
        rc = self.lib.mdb_infer_preprocessed_rp(self.h, ptrs, lh, lw, sh, sw, gp.ctypes.data if gp is not None else None, n,
                                                float(conf_thres), det.ctypes.data, cnt.ctypes.data,
                                                anc.ctypes.data if anc is not None else None, status.ctypes.data)
        self._check(rc, 'mdb_infer_preprocessed_rp')
        return det, cnt, anc, status

    def infer_jpeg_batch(self, blobs, conf_thres, want_anchors=False):
        """blobs: list of `bytes` (JPEG files).  Returns (det, count, anchors, status, src_hw [n,2])."""
        n = len(blobs)
        bufs = [np.frombuffer(b, np.uint8) if b is not None and len(b) else None for b in blobs]
        ptrs = (C.c_void_p * n)(*[(a.ctypes.data if a is not None else None) for a in bufs])
        lens = (C.c_size_t * n)(*[(a.size if a is not None else 0) for a in bufs])
        det = np.zeros((n, self.max_det, 6), np.float32)
        cnt = np.zeros((n,), np.int32)
        anc = np.full((n, self.max_det), -1, np.int32) if want_anchors else None
        status = np.zeros((n,), np.int32)
        hw = np.zeros((n, 2), np.int32)
        rc = self.lib.mdb_infer_jpeg_batch(self.h, ptrs, lens, n, float(conf_thres), det.ctypes.data, cnt.ctypes.data,
                                           anc.ctypes.data if anc is not None else None, status.ctypes.data, hw.ctypes.data)
        self._check(rc, 'mdb_infer_jpeg_batch')
        return det, cnt, anc, status, hw

    def op_decode_jpeg(self, blob):
        """-> (rgb [h,w,3] u8, backend) for one JPEG file's bytes."""
        a = np.frombuffer(blob, np.uint8)
        hw = (C.c_int32 * 2)()
        be = C.c_int32(0)
        self._check(self.lib.mdb_op_decode_jpeg(self.h, a.ctypes.data, a.size, None, 0, hw, C.byref(be)), 'mdb_op_decode_jpeg')
        out = np.empty((hw[0], hw[1], 3), np.uint8)
        self._check(self.lib.mdb_op_decode_jpeg(self.h, a.ctypes.data, a.size, out.ctypes.data, out.nbytes, hw, C.byref(be)),
                    'mdb_op_decode_jpeg')
        return out, int(be.value)

    def set_augment(self, on):
        self._check(self.lib.mdb_set_augment(self.h, int(bool(on))), 'mdb_set_augment')

    def set_image_size(self, img_size):
        self._check(self.lib.mdb_set_image_size(self.h, int(img_size)), 'mdb_set_image_size')
        self.cfg.img_size = int(img_size)

    def last_timing(self):
        ms = (C.c_float * 4)()
        self._check(self.lib.mdb_last_timing(self.h, ms), 'mdb_last_timing')
        return {'preprocess_ms': ms[0], 'network_ms': ms[1], 'postprocess_ms': ms[2], 'total_ms': ms[3]}

    def bench_resident(self, iters, conf_thres, stages_mask=7):
        """Device time (ms) of `iters` back-to-back passes over the batch left staged in HBM by the last infer call."""
        ms = C.c_float(0)
        self._check(self.lib.mdb_bench_resident(self.h, int(iters), float(conf_thres), int(stages_mask), C.byref(ms)),
                    'mdb_bench_resident')
        return float(ms.value)

    def profile_steps(self, iters=5):
        """[(description, ms, macs)] per plan step for the batch left staged by the last infer call."""
        cap, stride = 512, 160
        ms = np.zeros(cap, np.float32)
        macs = np.zeros(cap, np.float64)
        names = C.create_string_buffer(cap * stride)
        n = C.c_int32(0)
        self._check(self.lib.mdb_profile_steps(self.h, int(iters), ms.ctypes.data, macs.ctypes.data, names, stride, cap,
                                               C.byref(n)), 'mdb_profile_steps')
        out = []
        for i in range(min(n.value, cap)):
            out.append((names.raw[i * stride:(i + 1) * stride].split(b'\0', 1)[0].decode(), float(ms[i]), float(macs[i])))
        return out

    def set_pipeline_split(self, parts):
        self._check(self.lib.mdb_set_pipeline_split(self.h, int(parts)), 'mdb_set_pipeline_split')

    def launch_count(self):
        return int(self.lib.mdb_launch_count(self.h))

    def fetch(self, what, index=0):
        shp = (C.c_int64 * 4)()
        self._check(self.lib.mdb_debug_fetch(self.h, what, index, None, 0, shp), 'mdb_debug_fetch')
        shp = [int(x) for x in shp]
        if what == MDB_FETCH_STEM_INPUT:
            out = np.empty(shp, np.float16)
        elif what == MDB_FETCH_PRED:
            out = np.empty(shp[:3], np.float32)
        else:
            out = np.empty((shp[0], shp[1]), np.float16)
        self._check(self.lib.mdb_debug_fetch(self.h, what, index, out.ctypes.data, out.nbytes, shp_arr(shp)),
                    'mdb_debug_fetch')
        if what == MDB_FETCH_LAYER:
            return out[:, shp[2]:shp[2] + shp[3]]
        return out

    # ---- parity harness: per-step trace ----
    def set_trace(self, names):
        """names: None (off), '*' (every step) or an iterable of step names."""
        if names is None:
            arg = None
        elif isinstance(names, str):
            arg = names.encode()
        else:
            arg = ','.join(names).encode()
        self._check(self.lib.mdb_debug_set_trace(self.h, arg), 'mdb_debug_set_trace')

    def trace_names(self):
        buf = C.create_string_buffer(1 << 16)
        n = C.c_int32(0)
        self._check(self.lib.mdb_debug_trace_names(self.h, buf, len(buf), C.byref(n)), 'mdb_debug_trace_names')
        return [x for x in buf.value.decode().split('\n') if x]

    def fetch_trace(self, name, which):
        """which: 0 input, 1 output, 2 residual -> numpy [N,H,W,C] (fp16, or fp32 for the Detect heads' output) or None."""
        shp = (C.c_int64 * 5)()
        self._check(self.lib.mdb_debug_fetch_trace(self.h, name.encode(), which, None, 0, shp), 'mdb_debug_fetch_trace')
        if shp[0] == 0:
            return None
        out = np.empty([int(shp[i]) for i in range(4)], np.float32 if shp[4] else np.float16)
        self._check(self.lib.mdb_debug_fetch_trace(self.h, name.encode(), which, out.ctypes.data, out.nbytes, shp),
                    'mdb_debug_fetch_trace')
        return out

    # ---- single-kernel entry points (parity tests) ----
    def op_upsample2x(self, x):
        x = np.ascontiguousarray(x, np.float16)
        n, h, w, c = x.shape
        y = np.empty((n, 2 * h, 2 * w, c), np.float16)
        self._check(self.lib.mdb_op_upsample2x(self.h, x.ctypes.data, n, h, w, c, y.ctypes.data), 'mdb_op_upsample2x')
        return y

    def op_sppf_pool(self, s1):
        s1 = np.ascontiguousarray(s1, np.float16)
        n, h, w, c = s1.shape
        y = np.empty((n, h, w, 4 * c), np.float16)
        self._check(self.lib.mdb_op_sppf_pool(self.h, s1.ctypes.data, n, h, w, c, y.ctypes.data), 'mdb_op_sppf_pool')
        return y

    def op_decode(self, raws):
        """raws: four fp32 arrays [n][ny][nx][round_up(3*no, 4)] (P3..P6) -> pred [n][anchors][no]."""
        raws = [np.ascontiguousarray(r, np.float32) for r in raws]
        n = raws[0].shape[0]
        ptrs = (C.c_void_p * 4)(*[r.ctypes.data for r in raws])
        ny = (C.c_int32 * 4)(*[r.shape[1] for r in raws])
        nx = (C.c_int32 * 4)(*[r.shape[2] for r in raws])
        rows = sum(3 * r.shape[1] * r.shape[2] for r in raws)
        pred = np.empty((n, rows, self.no), np.float32)
        self._check(self.lib.mdb_op_decode(self.h, ptrs, ny, nx, n, pred.ctypes.data), 'mdb_op_decode')
        return pred

    def op_letterbox(self, img):
        g = letterbox_geometry(self.cfg, img.shape[0], img.shape[1])
        oh, ow = g['out_hw']
        out = np.empty((oh // 2, ow // 2, 48), np.float16)
        hw = (C.c_int32 * 2)()
        img = np.ascontiguousarray(img)
        self._check(self.lib.mdb_op_letterbox(self.h, img.ctypes.data, img.shape[0], img.shape[1], img.strides[0],
                                              out.ctypes.data, hw), 'mdb_op_letterbox')
        return out

    def op_conv2d(self, x, w, b, k, s, pad, act=True, residual=None, out_fp32=False, iters=1):
        """x: fp16 NHWC; w: fp32 [co][ci][k][k]; b: fp32 [co]; returns (y NHWC, ms)."""
        x = np.ascontiguousarray(x, np.float16)
        w = np.ascontiguousarray(w, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        n, h, wd, ci = x.shape
        co = w.shape[0]
        ho, wo = (h + 2 * pad - k) // s + 1, (wd + 2 * pad - k) // s + 1
        y = np.empty((n, ho, wo, co), np.float32 if out_fp32 else np.float16)
        r = None
        if residual is not None:
            r = np.ascontiguousarray(residual, np.float16)
        ms = C.c_float(0)
        self._check(self.lib.mdb_op_conv2d(self.h, x.ctypes.data, n, h, wd, ci, w.ctypes.data, b.ctypes.data, co, k, s,
                                           pad, int(bool(act)), r.ctypes.data if r is not None else None,
                                           int(bool(out_fp32)), y.ctypes.data, iters, C.byref(ms)), 'mdb_op_conv2d')
        return y, float(ms.value)

    def op_nms(self, pred, conf_thres, iou_thres=0.45, max_det=300):
        pred = np.ascontiguousarray(pred, np.float32)
        n, a, no = pred.shape
        det = np.zeros((n, max_det, 6), np.float32)
        cnt = np.zeros((n,), np.int32)
        anc = np.full((n, max_det), -1, np.int32)
        self._check(self.lib.mdb_op_nms(self.h, pred.ctypes.data, n, a, no - 5, float(conf_thres), float(iou_thres),
                                        max_det, det.ctypes.data, cnt.ctypes.data, anc.ctypes.data), 'mdb_op_nms')
        return det, cnt, anc


def shp_arr(shp):
    a = (C.c_int64 * 4)()
    for i in range(4):
        a[i] = shp[i]
    return a
