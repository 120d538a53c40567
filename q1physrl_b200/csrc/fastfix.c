/*
 * fastfix.c -- CPython extension `q1physrl_b200._fastfix`: the reference's ActionDecoder._fix_actions
 * (q1physrl_env/env.py:221-223) without the interpreter in the inner loop.
 *
 * RLLib hands VectorPhysEnv.vector_step a list of N per-env tuples whose nk+1 elements are Python /
 * NumPy scalars or 1-element arrays; the reference normalises them with a Python double loop and
 * np.ravel per element, which is 94 % of its vector_step wall time.  This
 * walks the same nested sequence in C and writes the (N, width) float64 array the decoder consumes:
 * element value = the first item of the flattened element, converted to double -- what
 * np.array([[np.ravel(x)[0] for x in a] for a in actions], dtype=float64) yields.
 *
 *   fix_actions(actions, width, out) -> None      out: writable C-contiguous float64 buffer (N*width)
 *
 * Raises (TypeError / ValueError) for anything it does not recognise; the caller then takes the
 * reference's own element-wise route, which produces the reference's own error behaviour.
 * Plain CPython API + the buffer protocol; no NumPy headers needed.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

/* first element of a buffer-protocol object (NumPy arrays and NumPy scalars), as double */
static int first_of_buffer(PyObject *obj, double *out)
{
    Py_buffer view;
    if (PyObject_GetBuffer(obj, &view, PyBUF_FORMAT | PyBUF_ND | PyBUF_STRIDES) != 0) {
        PyErr_Clear();
        return -1;
    }
    int rc = -1;
    if (view.len >= view.itemsize && view.itemsize > 0 && view.buf) {
        const char *f = view.format ? view.format : "B";
        while (*f == '@' || *f == '=' || *f == '<' || *f == '|')   /* native / little-endian only */
            f++;
        const void *p = view.buf;   /* row-major first element is at the base address */
        rc = 0;
        if (f[0] && f[1] == '\0') {
            switch (f[0]) {
            case 'd': { double v; memcpy(&v, p, 8); *out = v; break; }
            case 'f': { float v; memcpy(&v, p, 4); *out = (double)v; break; }
            case 'e': rc = -1; break;                                  /* float16: let NumPy do it */
            case '?': *out = *(const unsigned char *)p ? 1.0 : 0.0; break;
            case 'b': *out = (double)*(const signed char *)p; break;
            case 'B': *out = (double)*(const unsigned char *)p; break;
            case 'h': { int16_t v; memcpy(&v, p, 2); *out = (double)v; break; }
            case 'H': { uint16_t v; memcpy(&v, p, 2); *out = (double)v; break; }
            case 'i': { int32_t v; memcpy(&v, p, 4); *out = (double)v; break; }
            case 'I': { uint32_t v; memcpy(&v, p, 4); *out = (double)v; break; }
            case 'l': if (view.itemsize == 8) { int64_t v; memcpy(&v, p, 8); *out = (double)v; }
                      else { int32_t v; memcpy(&v, p, 4); *out = (double)v; } break;
            case 'L': if (view.itemsize == 8) { uint64_t v; memcpy(&v, p, 8); *out = (double)v; }
                      else { uint32_t v; memcpy(&v, p, 4); *out = (double)v; } break;
            case 'q': { int64_t v; memcpy(&v, p, 8); *out = (double)v; break; }
            case 'Q': { uint64_t v; memcpy(&v, p, 8); *out = (double)v; break; }
            default: rc = -1;
            }
        } else {
            rc = -1;
        }
    }
    PyBuffer_Release(&view);
    return rc;
}

static int element_value(PyObject *x, double *out)
{
    if (PyFloat_CheckExact(x)) {
        *out = PyFloat_AS_DOUBLE(x);
        return 0;
    }
    if (PyLong_CheckExact(x) || PyBool_Check(x)) {
        double v = PyLong_AsDouble(x);
        if (v == -1.0 && PyErr_Occurred())
            return -1;
        *out = v;
        return 0;
    }
    if (PyObject_CheckBuffer(x) && first_of_buffer(x, out) == 0)
        return 0;
    if (PyList_Check(x) || PyTuple_Check(x)) {       /* nested sequence: np.ravel(x)[0] is its first leaf */
        if (PySequence_Fast_GET_SIZE(x) < 1) {
            PyErr_SetString(PyExc_ValueError, "empty action element");
            return -1;
        }
        return element_value(PySequence_Fast_GET_ITEM(x, 0), out);
    }
    PyObject *f = PyNumber_Float(x);                  /* anything else that knows how to be a float */
    if (!f)
        return -1;
    *out = PyFloat_AS_DOUBLE(f);
    Py_DECREF(f);
    return 0;
}

static PyObject *fix_actions(PyObject *self, PyObject *args)
{
    PyObject *actions, *out_obj;
    Py_ssize_t width;
    if (!PyArg_ParseTuple(args, "OnO", &actions, &width, &out_obj))
        return NULL;
    PyObject *outer = PySequence_Fast(actions, "actions must be a sequence of per-env action tuples");
    if (!outer)
        return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(outer);
    Py_buffer out;
    if (PyObject_GetBuffer(out_obj, &out, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) {
        Py_DECREF(outer);
        return NULL;
    }
    PyObject *result = NULL;
    if (width < 1 || out.len != (Py_ssize_t)(n * width * sizeof(double))) {
        PyErr_SetString(PyExc_ValueError, "out must hold len(actions) * width float64 values");
        goto done;
    }
    double *dst = (double *)out.buf;
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *row = PySequence_Fast_GET_ITEM(outer, i);
        if (!(PyTuple_Check(row) || PyList_Check(row))) {
            PyErr_SetString(PyExc_TypeError, "each action must be a tuple or a list");
            goto done;
        }
        if (PySequence_Fast_GET_SIZE(row) != width) {
            PyErr_SetString(PyExc_ValueError, "action with an unexpected number of elements");
            goto done;
        }
        for (Py_ssize_t j = 0; j < width; j++)
            if (element_value(PySequence_Fast_GET_ITEM(row, j), dst + i * width + j) != 0) {
                if (!PyErr_Occurred())
                    PyErr_SetString(PyExc_TypeError, "unsupported action element");
                goto done;
            }
    }
    result = Py_None;
    Py_INCREF(result);
done:
    PyBuffer_Release(&out);
    Py_DECREF(outer);
    return result;
}

/* split_actions(actions, num_keys, has_mouse, keys_out, mouse_out) -> None
 * The same walk, written straight into the two arrays q1_step_host consumes: keys_out (N, num_keys)
 * uint8 = bit 0 of the key action truncated to an integer (env:228 `.astype(np.int)`, env:243 `&`
 * against 0/1 values), mouse_out (N,) float64 (ignored when has_mouse is 0). */
static PyObject *split_actions(PyObject *self, PyObject *args)
{
    PyObject *actions, *keys_obj, *mouse_obj;
    Py_ssize_t nk;
    int has_mouse;
    if (!PyArg_ParseTuple(args, "OnpOO", &actions, &nk, &has_mouse, &keys_obj, &mouse_obj))
        return NULL;
    PyObject *outer = PySequence_Fast(actions, "actions must be a sequence of per-env action tuples");
    if (!outer)
        return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(outer);
    const Py_ssize_t width = nk + (has_mouse ? 1 : 0);
    Py_buffer kb, mb;
    mb.buf = NULL;
    if (PyObject_GetBuffer(keys_obj, &kb, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) {
        Py_DECREF(outer);
        return NULL;
    }
    if (has_mouse && PyObject_GetBuffer(mouse_obj, &mb, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) {
        PyBuffer_Release(&kb);
        Py_DECREF(outer);
        return NULL;
    }
    PyObject *result = NULL;
    if (nk < 1 || kb.len != n * nk || (has_mouse && mb.len != (Py_ssize_t)(n * sizeof(double)))) {
        PyErr_SetString(PyExc_ValueError, "keys_out / mouse_out do not match len(actions)");
        goto done;
    }
    unsigned char *keys = (unsigned char *)kb.buf;
    double *mouse = has_mouse ? (double *)mb.buf : NULL;
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *row = PySequence_Fast_GET_ITEM(outer, i);
        if (!(PyTuple_Check(row) || PyList_Check(row))) {
            PyErr_SetString(PyExc_TypeError, "each action must be a tuple or a list");
            goto done;
        }
        if (PySequence_Fast_GET_SIZE(row) < width) {
            PyErr_SetString(PyExc_ValueError, "action with too few elements");
            goto done;
        }
        for (Py_ssize_t j = 0; j < width; j++) {
            double v;
            if (element_value(PySequence_Fast_GET_ITEM(row, j), &v) != 0) {
                if (!PyErr_Occurred())
                    PyErr_SetString(PyExc_TypeError, "unsupported action element");
                goto done;
            }
            if (j < nk) {
                if (!(v > -9.2e18 && v < 9.2e18)) {      /* astype(int) of nan / inf / huge: NumPy's business */
                    PyErr_SetString(PyExc_ValueError, "key action outside the int64 range");
                    goto done;
                }
                keys[i * nk + j] = (unsigned char)((long long)v & 1);
            } else {
                mouse[i] = v;
            }
        }
    }
    result = Py_None;
    Py_INCREF(result);
done:
    if (mb.buf)
        PyBuffer_Release(&mb);
    PyBuffer_Release(&kb);
    Py_DECREF(outer);
    return result;
}

/* step(fn, handle, keys, mouse | None, mouse_kind, obs, reward, done, zero_start | None, auto_reset)
 * -> the int q1_step_host returns.  `fn` is the address of q1_step_host (include/q1phys.h) and
 * `handle` the q1_env*, both as integers.  Takes the array addresses through the buffer protocol
 * and makes the C call directly: at RLLib's 100 envs per worker the ctypes marshalling of nine
 * arguments and seven `ndarray.ctypes` objects costs as much as the GPU work. */
typedef int (*step_host_fn)(void *, const void *, const void *, int, void *, void *, void *, void *, int);

static PyObject *step(PyObject *self, PyObject *args)
{
    unsigned long long fn, handle;
    PyObject *objs[6];   /* keys, mouse, obs, reward, done, zero_start */
    int mouse_kind, auto_reset;
    if (!PyArg_ParseTuple(args, "KKOOiOOOOp", &fn, &handle, &objs[0], &objs[1], &mouse_kind, &objs[2],
                          &objs[3], &objs[4], &objs[5], &auto_reset))
        return NULL;
    static const int writable[6] = {0, 0, 1, 1, 1, 1};
    Py_buffer view[6];
    void *ptr[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
    int got = 0;
    for (; got < 6; got++) {
        view[got].buf = NULL;
        if (objs[got] == Py_None)
            continue;
        if (PyObject_GetBuffer(objs[got], &view[got],
                               (writable[got] ? PyBUF_WRITABLE : 0) | PyBUF_C_CONTIGUOUS) != 0)
            break;
        ptr[got] = view[got].buf;
    }
    PyObject *result = NULL;
    if (got == 6) {
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = ((step_host_fn)(uintptr_t)fn)((void *)(uintptr_t)handle, ptr[0], ptr[1], mouse_kind, ptr[2],
                                           ptr[3], ptr[4], ptr[5], auto_reset);
        Py_END_ALLOW_THREADS
        result = PyLong_FromLong(rc);
    }
    for (int k = 0; k < got; k++)
        if (view[k].buf)
            PyBuffer_Release(&view[k]);
    return result;
}

/* step_arrays(fn, handle, n, num_keys, has_mouse, keys, mouse, obs, reward, done, zero_start, auto_reset)
 * -> rc of q1_step_host, or -100 when the action arrays are not in a layout the library takes as it
 * is (the caller then normalises them in Python and calls `step`).  Accepted: keys C-contiguous, 1-byte
 * integer / bool items, n * num_keys of them; mouse C-contiguous n items of float32 ('f'), int32 ('i')
 * or float64 ('d'); the element type selects Q1_MOUSE_*.  The checks a Python caller would make on
 * dtype / shape / contiguity are made here on the buffer views instead. */
static PyObject *step_arrays(PyObject *self, PyObject *args)
{
    unsigned long long fn, handle;
    Py_ssize_t n, nk;
    int has_mouse, auto_reset;
    PyObject *keys_obj, *mouse_obj, *out_obj[4];
    if (!PyArg_ParseTuple(args, "KKnnpOOOOOOp", &fn, &handle, &n, &nk, &has_mouse, &keys_obj, &mouse_obj,
                          &out_obj[0], &out_obj[1], &out_obj[2], &out_obj[3], &auto_reset))
        return NULL;
    Py_buffer kb, mb, ob[4];
    if (PyObject_GetBuffer(keys_obj, &kb, PyBUF_FORMAT | PyBUF_C_CONTIGUOUS) != 0) {
        PyErr_Clear();
        return PyLong_FromLong(-100);
    }
    long rc = -100;
    int have_mouse = 0, outs = 0, mouse_kind = 0;
    const char *kf = kb.format ? kb.format : "B";
    if (kb.itemsize != 1 || kb.len != n * nk || !(kf[0] == 'B' || kf[0] == 'b' || kf[0] == '?') ||
        kb.ndim != 2)
        goto done;
    if (has_mouse) {
        if (PyObject_GetBuffer(mouse_obj, &mb, PyBUF_FORMAT | PyBUF_C_CONTIGUOUS) != 0) {
            PyErr_Clear();
            goto done;
        }
        have_mouse = 1;
        const char *mf = mb.format ? mb.format : "";
        while (*mf == '@' || *mf == '=' || *mf == '<')
            mf++;
        if (mb.ndim != 1 || mb.shape[0] != n || mf[1] != '\0')
            goto done;
        if (mf[0] == 'f' && mb.itemsize == 4)
            mouse_kind = 0;                                      /* Q1_MOUSE_F32 */
        else if (mf[0] == 'i' && mb.itemsize == 4)
            mouse_kind = 1;                                      /* Q1_MOUSE_I32 */
        else if (mf[0] == 'd' && mb.itemsize == 8)
            mouse_kind = 2;                                      /* Q1_MOUSE_F64 */
        else
            goto done;
    }
    for (; outs < 4; outs++)
        if (PyObject_GetBuffer(out_obj[outs], &ob[outs], PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0)
            break;
    if (outs < 4) {
        for (int k = 0; k < outs; k++)
            PyBuffer_Release(&ob[k]);
        if (have_mouse)
            PyBuffer_Release(&mb);
        PyBuffer_Release(&kb);
        return NULL;
    }
    if (ob[0].len != n * 24 || ob[1].len != n * 4 || ob[2].len != n || ob[3].len != n) {
        PyErr_SetString(PyExc_ValueError, "output arrays do not match num_envs");
        rc = -101;
        goto done;
    }
    Py_BEGIN_ALLOW_THREADS
    rc = ((step_host_fn)(uintptr_t)fn)((void *)(uintptr_t)handle, kb.buf, have_mouse ? mb.buf : NULL, mouse_kind,
                                       ob[0].buf, ob[1].buf, ob[2].buf, ob[3].buf, auto_reset);
    Py_END_ALLOW_THREADS
done:
    for (int k = 0; k < outs; k++)
        PyBuffer_Release(&ob[k]);
    if (have_mouse)
        PyBuffer_Release(&mb);
    PyBuffer_Release(&kb);
    if (rc == -101)
        return NULL;
    return PyLong_FromLong(rc);
}

static PyMethodDef methods[] = {
    {"step_arrays", step_arrays, METH_VARARGS,
     "step_arrays(fn, handle, n, num_keys, has_mouse, keys, mouse, obs, reward, done, zero_start, auto_reset) "
     "-> rc of q1_step_host, or -100 if the action arrays need normalising first."},
    {"split_actions", split_actions, METH_VARARGS,
     "split_actions(actions, num_keys, has_mouse, keys_out, mouse_out): RLLib's nested action format -> "
     "the uint8 key array and float64 mouse array q1_step_host consumes (reference env.py:221-228)."},
    {"step", step, METH_VARARGS,
     "step(fn, handle, keys, mouse, mouse_kind, obs, reward, done, zero_start, auto_reset) -> rc of "
     "q1_step_host called through its address."},
    {"fix_actions", fix_actions, METH_VARARGS,
     "fix_actions(actions, width, out): normalise RLLib's nested action format into the (N, width) "
     "float64 buffer `out` (reference env.py:221-223)."},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fastfix", NULL, -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__fastfix(void) { return PyModule_Create(&module); }
