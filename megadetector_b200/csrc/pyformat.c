// Host-side record formatting as a CPython extension (`megadetector_b200._mdbfmt`): the per-detection loop of
// /root/reference/megadetector/detection/pytorch_detector.py:1384-1411 -- xyxy2xywh and `/ gn` in float32, convert_yolo_to_xywh
// (utils/ct_utils.py:255-270) in double, truncate_float (:82-103) or round_float (Python round()), category string, dict -- for
// one image per call.  The numpy version in detector.py (format_detections_numpy) states the same arithmetic and is what the
// tests compare this against; the engine's last chunk of a batch has nothing left to overlap its formatting with, so the
// ~0.65 us per detection of the numpy version (300 detections x 16 images = 3 ms) were exposed in every end-to-end step.
//
// Built by megadetector_b200/build.py with gcc -O2 -ffp-contract=off (no FMA contraction: the float32 steps must round where
// numpy's do).
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <stdint.h>

static PyObject* g_key_category;
static PyObject* g_key_conf;
static PyObject* g_key_bbox;
static PyObject* g_cat_str[16];  // str(c) for small class ids

static double truncate_d(double x, double f) { return floor(x * f) / f; }

// Python round(x, nd) for nd > 0: correctly rounded decimal, half to even -- what float.__round__ does (Objects/floatobject.c
// double_round): shortest 'f' formatting with nd digits through dtoa, parsed back (the non-classic modes only)
static int py_round(double x, int nd, double* out) {
  if (!isfinite(x)) { *out = x; return 0; }
  char* buf = PyOS_double_to_string(x, 'f', nd, 0, NULL);
  if (!buf) return -1;
  const double r = PyOS_string_to_double(buf, NULL, NULL);
  PyMem_Free(buf);
  if (r == -1.0 && PyErr_Occurred()) return -1;
  *out = r;
  return 0;
}

// format_detections(det, count, src_h, src_w, use_model_native_classes, classic) -> (list of dicts ascending conf, max_conf)
static PyObject* format_detections(PyObject* self, PyObject* args) {
  Py_buffer view;
  long count, src_h, src_w;
  int native_classes, classic;
  if (!PyArg_ParseTuple(args, "y*lllpp", &view, &count, &src_h, &src_w, &native_classes, &classic)) return NULL;
  PyObject* result = NULL;
  PyObject* list = NULL;
  if (count < 0 || (Py_ssize_t)(count * 6 * sizeof(float)) > view.len) {
    PyErr_SetString(PyExc_ValueError, "det buffer holds fewer than count rows of 6 float32");
    goto done;
  }
  if (count == 0) {
    result = Py_BuildValue("([]d)", 0.0);
    goto done;
  }
  {
    const float* det = (const float*)view.buf;
    const float gw = (float)src_w, gh = (float)src_h;  // gn = [W, H, W, H] (int64 -> float32 in the division)
    const double fc = 1e4, fs = 1e3;                    // COORD_DIGITS = 4, CONF_DIGITS = 3
    double max_conf = 0.0;
    list = PyList_New(count);
    if (!list) goto done;
    for (long i = 0; i < count; ++i) {
      const float* r = det + (count - 1 - i) * 6;  // `for *xyxy, conf, cls in reversed(det)` (:1384)
      const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
      volatile float sx = x1 + x2, sy = y1 + y2;  // volatile: every float32 step is rounded to float32, as numpy / torch do
      volatile float hx = sx / 2.0f, hy = sy / 2.0f;
      volatile float xc = hx / gw, yc = hy / gh;
      volatile float dw = x2 - x1, dh = y2 - y1;
      volatile float bw = dw / gw, bh = dh / gh;
      double v[4];
      v[2] = (double)bw;
      v[3] = (double)bh;
      v[0] = (double)xc - v[2] / 2.0;
      v[1] = (double)yc - v[3] / 2.0;
      double conf = (double)r[4];
      if (classic) {
        for (int k = 0; k < 4; ++k) v[k] = truncate_d(v[k], fc);
        conf = truncate_d(conf, fs);
      } else {
        for (int k = 0; k < 4; ++k)
          if (py_round(v[k], 4, &v[k]) < 0) goto fail;
        if (py_round(conf, 3, &conf) < 0) goto fail;
      }
      long cls = (long)r[5];  // astype(int64): truncation toward zero
      if (!native_classes) {
        cls += 1;
        if (cls < 1 || cls > 3) {
          PyObject* msg = PyUnicode_FromFormat("%ld is not a valid class.", cls);
          if (msg) { PyErr_SetObject(PyExc_KeyError, msg); Py_DECREF(msg); }
          goto fail;
        }
      }
      PyObject* cat;
      if (cls >= 0 && cls < 16) { cat = g_cat_str[cls]; Py_INCREF(cat); }
      else { cat = PyUnicode_FromFormat("%ld", cls); if (!cat) goto fail; }
      PyObject* bbox = PyList_New(4);
      PyObject* cf = PyFloat_FromDouble(conf);
      PyObject* d = PyDict_New();
      if (!bbox || !cf || !d) { Py_XDECREF(bbox); Py_XDECREF(cf); Py_XDECREF(d); Py_DECREF(cat); goto fail; }
      int bad = 0;
      for (int k = 0; k < 4; ++k) {
        PyObject* f = PyFloat_FromDouble(v[k]);
        if (!f) { bad = 1; break; }
        PyList_SET_ITEM(bbox, k, f);
      }
      // insertion order category, conf, bbox as in the reference's dict literal (:1404-1408)
      if (bad || PyDict_SetItem(d, g_key_category, cat) < 0 || PyDict_SetItem(d, g_key_conf, cf) < 0 ||
          PyDict_SetItem(d, g_key_bbox, bbox) < 0) {
        Py_DECREF(bbox); Py_DECREF(cf); Py_DECREF(d); Py_DECREF(cat);
        goto fail;
      }
      Py_DECREF(bbox); Py_DECREF(cf); Py_DECREF(cat);
      PyList_SET_ITEM(list, i, d);
      if (conf > max_conf) max_conf = conf;
    }
    result = Py_BuildValue("(Nd)", list, max_conf);
    list = NULL;
    goto done;
  }
fail:
  Py_XDECREF(list);
  list = NULL;
done:
  PyBuffer_Release(&view);
  return result;
}

static PyMethodDef methods[] = {
    {"format_detections", format_detections, METH_VARARGS,
     "format_detections(det_f32[n][6], count, src_h, src_w, use_model_native_classes, classic) -> (list of dicts, max_conf)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_mdbfmt", "MegaDetector record formatting (host)", -1, methods};

PyMODINIT_FUNC PyInit__mdbfmt(void) {
  PyObject* m = PyModule_Create(&moduledef);
  if (!m) return NULL;
  g_key_category = PyUnicode_InternFromString("category");
  g_key_conf = PyUnicode_InternFromString("conf");
  g_key_bbox = PyUnicode_InternFromString("bbox");
  for (int i = 0; i < 16; ++i) g_cat_str[i] = PyUnicode_FromFormat("%d", i);
  return m;
}
