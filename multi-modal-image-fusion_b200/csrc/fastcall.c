/* CPython binding of the two entries a training step calls every iteration (mmif_fusion_loss_fwd, mmif_fusion_loss_bwd3):
 * METH_FASTCALL wrappers with the exact C prototypes of include/mmif_b200.h.  ctypes spends ~9 us converting the 12-15
 * arguments of one call (tools/host_profile.py); this path spends ~0.5 us.  Same library, same entry points: the ctypes
 * binding in _lib.py stays the reference binding and the fallback when this module was not built. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include "../../include/mmif_b200.h"

static int as_ptr(PyObject* o, void** out) {
    if (o == Py_None) { *out = NULL; return 0; }
    unsigned long long v = PyLong_AsUnsignedLongLong(o);
    if (v == (unsigned long long)-1 && PyErr_Occurred()) return -1;
    *out = (void*)(uintptr_t)v;
    return 0;
}
static int as_int(PyObject* o, int* out) {
    long v = PyLong_AsLong(o);
    if (v == -1 && PyErr_Occurred()) return -1;
    *out = (int)v;
    return 0;
}
static int as_size(PyObject* o, size_t* out) {
    unsigned long long v = PyLong_AsUnsignedLongLong(o);
    if (v == (unsigned long long)-1 && PyErr_Occurred()) return -1;
    *out = (size_t)v;
    return 0;
}

/* loss_fwd(i1, i2, f, B, H, W, cfg_addr, out, dF_unit, ws, ws_bytes, stream) -> rc */
static PyObject* loss_fwd(PyObject* self, PyObject* const* a, Py_ssize_t n) {
    void *i1, *i2, *f, *cfg, *out, *dfu, *ws, *st;
    int B, H, W;
    size_t wsb;
    if (n != 12) { PyErr_SetString(PyExc_TypeError, "loss_fwd takes 12 arguments"); return NULL; }
    if (as_ptr(a[0], &i1) || as_ptr(a[1], &i2) || as_ptr(a[2], &f) || as_int(a[3], &B) || as_int(a[4], &H) || as_int(a[5], &W) ||
        as_ptr(a[6], &cfg) || as_ptr(a[7], &out) || as_ptr(a[8], &dfu) || as_ptr(a[9], &ws) || as_size(a[10], &wsb) || as_ptr(a[11], &st))
        return NULL;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = mmif_fusion_loss_fwd((const float*)i1, (const float*)i2, (const float*)f, B, H, W, (const MmifLossCfg*)cfg, (double*)out,
                              (float*)dfu, ws, wsb, st);
    Py_END_ALLOW_THREADS
    return PyLong_FromLong(rc);
}

/* loss_bwd3(i1, i2, f, B, H, W, cfg_addr, g_ssim, g_pixel, g_grad, dF_unit, dF, ws, ws_bytes, stream) -> rc */
static PyObject* loss_bwd3(PyObject* self, PyObject* const* a, Py_ssize_t n) {
    void *i1, *i2, *f, *cfg, *g0, *g1, *g2, *dfu, *df, *ws, *st;
    int B, H, W;
    size_t wsb;
    if (n != 15) { PyErr_SetString(PyExc_TypeError, "loss_bwd3 takes 15 arguments"); return NULL; }
    if (as_ptr(a[0], &i1) || as_ptr(a[1], &i2) || as_ptr(a[2], &f) || as_int(a[3], &B) || as_int(a[4], &H) || as_int(a[5], &W) ||
        as_ptr(a[6], &cfg) || as_ptr(a[7], &g0) || as_ptr(a[8], &g1) || as_ptr(a[9], &g2) || as_ptr(a[10], &dfu) || as_ptr(a[11], &df) ||
        as_ptr(a[12], &ws) || as_size(a[13], &wsb) || as_ptr(a[14], &st))
        return NULL;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = mmif_fusion_loss_bwd3((const float*)i1, (const float*)i2, (const float*)f, B, H, W, (const MmifLossCfg*)cfg, (const float*)g0,
                               (const float*)g1, (const float*)g2, (const float*)dfu, (float*)df, ws, wsb, st);
    Py_END_ALLOW_THREADS
    return PyLong_FromLong(rc);
}

static PyMethodDef methods[] = {
    {"loss_fwd", (PyCFunction)(void (*)(void))loss_fwd, METH_FASTCALL, "mmif_fusion_loss_fwd with integer addresses"},
    {"loss_bwd3", (PyCFunction)(void (*)(void))loss_bwd3, METH_FASTCALL, "mmif_fusion_loss_bwd3 with integer addresses"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastcall", "fast CPython binding of the per-step libmmif_b200 entries", -1, methods};
PyMODINIT_FUNC PyInit__fastcall(void) { return PyModule_Create(&moddef); }
