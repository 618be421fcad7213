/*
 * _mdf_pyhost: CPython glue between Python lists and the ragged entry points of libmdf_b200.
 *
 * The reference's pipeline holds, per alignment, a few Python strings and one small NumPy array (pipeline.py:476-481,
 * alignment.py:65-150).  Joining / concatenating 16k of them in Python costs more than the GPU needs for the whole path
 * (measured: 115 ms of packing against a 68 ms step), so this module only collects the POINTERS - the UTF-8 buffer of each str,
 * the data pointer of each float32 [rows, 3] array - and hands them to mdf_path_submit_ragged, which packs them into pinned
 * memory on worker threads with the GIL released.  No arithmetic happens here.
 *
 *   submit_ragged(fn_addr, model_addr, seqs, gapped_query, gapped_target, coords, thr2, gen, scores_addr) -> (rc, job_addr)
 *   pointers(strings) -> (bytes ptr-array, bytes len-array)      [used for the alignment / ingest entry points]
 *   cmap_ragged(fn_addr, ctx_addr, gapped_query, gapped_target, coords, thr2, gen, out_addr, out_words) -> (rc, packed_off, seq_off)
 *
 * Errors: a str that is not ASCII raises UnicodeEncodeError like predict.pyx:19; a coordinate item that is not a C-contiguous
 * float32 [rows, 3] buffer raises TypeError naming the item (the Python wrapper converts and retries).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct mdf_model mdf_model;
typedef struct mdf_job mdf_job;
typedef int (*submit_ragged_fn)(mdf_model *, int, const char *const *, const int *, const float *const *, const int *,
                                const char *const *, const char *const *, const int *, float, int, float *, mdf_job **);

/* borrowed pointer to the bytes of item `o` (str: ASCII only; bytes accepted) */
static int text_of(PyObject *o, const char **ptr, Py_ssize_t *len, const char *what, Py_ssize_t i)
{
    if (PyUnicode_Check(o)) {
        if (!PyUnicode_IS_ASCII(o)) {
            PyObject *b = PyUnicode_AsASCIIString(o);      /* raises UnicodeEncodeError with the offending position */
            Py_XDECREF(b);
            if (!PyErr_Occurred()) PyErr_Format(PyExc_ValueError, "%s[%zd] is not ASCII", what, i);
            return -1;
        }
        *ptr = PyUnicode_AsUTF8AndSize(o, len);            /* for ASCII strings: the string's own buffer, no copy */
        return *ptr ? 0 : -1;
    }
    if (PyBytes_Check(o)) {
        *ptr = PyBytes_AS_STRING(o);
        *len = PyBytes_GET_SIZE(o);
        return 0;
    }
    PyErr_Format(PyExc_TypeError, "%s[%zd] must be str or bytes", what, i);
    return -1;
}

static PyObject *submit_ragged(PyObject *self, PyObject *args)
{
    unsigned long long fn_addr, model_addr, scores_addr;
    PyObject *seqs, *gq, *gt, *coords;
    double thr2;
    int gen;
    if (!PyArg_ParseTuple(args, "KKOOOOdiK", &fn_addr, &model_addr, &seqs, &gq, &gt, &coords, &thr2, &gen, &scores_addr)) return NULL;
    PyObject *fs = PySequence_Fast(seqs, "seqs must be a sequence"), *fq = NULL, *ft = NULL, *fc = NULL;
    if (!fs) return NULL;
    fq = PySequence_Fast(gq, "gapped_query must be a sequence");
    ft = fq ? PySequence_Fast(gt, "gapped_target must be a sequence") : NULL;
    fc = ft ? PySequence_Fast(coords, "coords must be a sequence") : NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fs);
    const char **pseq = NULL, **pq = NULL, **pt = NULL;
    const float **pc = NULL;
    int *lseq = NULL, *laln = NULL, *rows = NULL;
    Py_buffer *views = NULL;
    Py_ssize_t nviews = 0;
    PyObject *result = NULL;
    if (!fc) goto done;
    if (PySequence_Fast_GET_SIZE(fq) != n || PySequence_Fast_GET_SIZE(ft) != n || PySequence_Fast_GET_SIZE(fc) != n) {
        PyErr_SetString(PyExc_ValueError, "seqs, gapped_query, gapped_target and coords must have the same length");
        goto done;
    }
    pseq = malloc(sizeof(char *) * (n + 1)); pq = malloc(sizeof(char *) * (n + 1)); pt = malloc(sizeof(char *) * (n + 1));
    pc = malloc(sizeof(float *) * (n + 1));
    lseq = malloc(sizeof(int) * (n + 1)); laln = malloc(sizeof(int) * (n + 1)); rows = malloc(sizeof(int) * (n + 1));
    views = calloc((size_t)n + 1, sizeof(Py_buffer));
    if (!pseq || !pq || !pt || !pc || !lseq || !laln || !rows || !views) { PyErr_NoMemory(); goto done; }
    for (Py_ssize_t i = 0; i < n; ++i) {
        Py_ssize_t ls, lq, lt;
        if (text_of(PySequence_Fast_GET_ITEM(fs, i), &pseq[i], &ls, "seqs", i) < 0) goto done;
        if (text_of(PySequence_Fast_GET_ITEM(fq, i), &pq[i], &lq, "gapped_query", i) < 0) goto done;
        if (text_of(PySequence_Fast_GET_ITEM(ft, i), &pt[i], &lt, "gapped_target", i) < 0) goto done;
        if (lq != lt) { PyErr_Format(PyExc_ValueError, "alignment %zd: query and target alignments differ in length", i); goto done; }
        if (ls > INT32_MAX || lq > INT32_MAX) { PyErr_SetString(PyExc_OverflowError, "sequence too long"); goto done; }
        lseq[i] = (int)ls; laln[i] = (int)lq;
        PyObject *c = PySequence_Fast_GET_ITEM(fc, i);
        Py_buffer *v = &views[nviews];
        if (PyObject_GetBuffer(c, v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) {
            PyErr_Clear();
            PyErr_Format(PyExc_TypeError, "coords[%zd] is not a C-contiguous buffer", i);
            goto done;
        }
        ++nviews;
        const char *f = v->format ? v->format : "";
        if (*f == '<' || *f == '=' || *f == '@') ++f;
        if (strcmp(f, "f") != 0 || v->itemsize != 4 || v->ndim != 2 || v->shape[1] != 3) {
            PyErr_Format(PyExc_TypeError, "coords[%zd] is not a float32 array of shape (Lt, 3)", i);
            goto done;
        }
        pc[i] = (const float *)v->buf;
        rows[i] = (int)v->shape[0];
    }
    {
        mdf_job *job = NULL;
        int rc;
        submit_ragged_fn fn = (submit_ragged_fn)(uintptr_t)fn_addr;
        Py_BEGIN_ALLOW_THREADS
        rc = fn((mdf_model *)(uintptr_t)model_addr, (int)n, pseq, lseq, pc, rows, pq, pt, laln, (float)thr2, gen,
                (float *)(uintptr_t)scores_addr, &job);
        Py_END_ALLOW_THREADS
        result = Py_BuildValue("iK", rc, (unsigned long long)(uintptr_t)job);
    }
done:
    for (Py_ssize_t i = 0; i < nviews; ++i) PyBuffer_Release(&views[i]);
    free(views); free(pseq); free(pq); free(pt); free(pc); free(lseq); free(laln); free(rows);
    Py_XDECREF(fs); Py_XDECREF(fq); Py_XDECREF(ft); Py_XDECREF(fc);
    return result;
}

/* pointers(list of str / bytes) -> (ptr-array as bytes [n x 8], int32 length array as bytes [n x 4]); the pointers stay valid as
 * long as the list's items live */
static PyObject *pointers(PyObject *self, PyObject *arg)
{
    PyObject *f = PySequence_Fast(arg, "expected a sequence of str / bytes");
    if (!f) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(f);
    PyObject *pb = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(char *)), *lb = PyBytes_FromStringAndSize(NULL, n * 4);
    if (!pb || !lb) { Py_XDECREF(pb); Py_XDECREF(lb); Py_DECREF(f); return NULL; }
    const char **pp = (const char **)PyBytes_AS_STRING(pb);
    int *ll = (int *)PyBytes_AS_STRING(lb);
    for (Py_ssize_t i = 0; i < n; ++i) {
        Py_ssize_t len;
        if (text_of(PySequence_Fast_GET_ITEM(f, i), &pp[i], &len, "item", i) < 0 || len > INT32_MAX) {
            if (!PyErr_Occurred()) PyErr_SetString(PyExc_OverflowError, "string too long");
            Py_DECREF(pb); Py_DECREF(lb); Py_DECREF(f);
            return NULL;
        }
        ll[i] = (int)len;
    }
    Py_DECREF(f);
    return Py_BuildValue("NN", pb, lb);
}

typedef struct mdf_ctx mdf_ctx;
typedef int (*cmap_ragged_fn)(mdf_ctx *, int, const float *const *, const int *, const char *const *, const char *const *, const int *,
                              float, int, uint32_t *, size_t, int64_t *, int64_t *);

/* cmap_ragged(fn_addr, ctx_addr, gapped_query, gapped_target, coords, thr2, gen, out_addr, out_capacity_words)
 *   -> (rc, packed_off bytes [n+1 x int64], seq_off bytes [n+1 x int64])
 * out_addr == 0: offsets only (coords / gapped_target may be None then).  Calls mdf_cmap_build_transfer_ragged with the GIL
 * released. */
static PyObject *cmap_ragged(PyObject *self, PyObject *args)
{
    unsigned long long fn_addr, ctx_addr, out_addr, out_cap;
    PyObject *gq, *gt, *coords;
    double thr2;
    int gen;
    if (!PyArg_ParseTuple(args, "KKOOOdiKK", &fn_addr, &ctx_addr, &gq, &gt, &coords, &thr2, &gen, &out_addr, &out_cap)) return NULL;
    const int sizes_only = out_addr == 0;
    PyObject *fq = PySequence_Fast(gq, "gapped_query must be a sequence"), *ft = NULL, *fc = NULL;
    if (!fq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fq);
    const char **pq = NULL, **pt = NULL;
    const float **pc = NULL;
    int *laln = NULL, *rows = NULL;
    Py_buffer *views = NULL;
    Py_ssize_t nviews = 0;
    PyObject *result = NULL, *poff = NULL, *soff = NULL;
    if (!sizes_only) {
        ft = PySequence_Fast(gt, "gapped_target must be a sequence");
        fc = ft ? PySequence_Fast(coords, "coords must be a sequence") : NULL;
        if (!fc) goto done;
        if (PySequence_Fast_GET_SIZE(ft) != n || PySequence_Fast_GET_SIZE(fc) != n) {
            PyErr_SetString(PyExc_ValueError, "gapped_query, gapped_target and coords must have the same length");
            goto done;
        }
    }
    pq = malloc(sizeof(char *) * (n + 1)); pt = malloc(sizeof(char *) * (n + 1)); pc = malloc(sizeof(float *) * (n + 1));
    laln = malloc(sizeof(int) * (n + 1)); rows = malloc(sizeof(int) * (n + 1));
    views = calloc((size_t)n + 1, sizeof(Py_buffer));
    poff = PyBytes_FromStringAndSize(NULL, (n + 1) * 8);
    soff = PyBytes_FromStringAndSize(NULL, (n + 1) * 8);
    if (!pq || !pt || !pc || !laln || !rows || !views || !poff || !soff) { PyErr_NoMemory(); goto done; }
    for (Py_ssize_t i = 0; i < n; ++i) {
        Py_ssize_t lq, lt;
        if (text_of(PySequence_Fast_GET_ITEM(fq, i), &pq[i], &lq, "gapped_query", i) < 0) goto done;
        if (lq > INT32_MAX) { PyErr_SetString(PyExc_OverflowError, "alignment too long"); goto done; }
        laln[i] = (int)lq;
        pt[i] = NULL; pc[i] = NULL; rows[i] = 0;
        if (sizes_only) continue;
        if (text_of(PySequence_Fast_GET_ITEM(ft, i), &pt[i], &lt, "gapped_target", i) < 0) goto done;
        if (lq != lt) { PyErr_Format(PyExc_ValueError, "alignment %zd: query and target alignments differ in length", i); goto done; }
        PyObject *c = PySequence_Fast_GET_ITEM(fc, i);
        Py_buffer *v = &views[nviews];
        if (PyObject_GetBuffer(c, v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) {
            PyErr_Clear();
            PyErr_Format(PyExc_TypeError, "coords[%zd] is not a C-contiguous buffer", i);
            goto done;
        }
        ++nviews;
        const char *f = v->format ? v->format : "";
        if (*f == '<' || *f == '=' || *f == '@') ++f;
        if (strcmp(f, "f") != 0 || v->itemsize != 4 || v->ndim != 2 || v->shape[1] != 3) {
            PyErr_Format(PyExc_TypeError, "coords[%zd] is not a float32 array of shape (Lt, 3)", i);
            goto done;
        }
        pc[i] = (const float *)v->buf;
        rows[i] = (int)v->shape[0];
    }
    {
        int rc;
        cmap_ragged_fn fn = (cmap_ragged_fn)(uintptr_t)fn_addr;
        int64_t *po = (int64_t *)PyBytes_AS_STRING(poff), *so = (int64_t *)PyBytes_AS_STRING(soff);
        Py_BEGIN_ALLOW_THREADS
        rc = fn((mdf_ctx *)(uintptr_t)ctx_addr, (int)n, pc, rows, pq, pt, laln, (float)thr2, gen, (uint32_t *)(uintptr_t)out_addr,
                (size_t)out_cap, po, so);
        Py_END_ALLOW_THREADS
        result = Py_BuildValue("iOO", rc, poff, soff);
    }
done:
    for (Py_ssize_t i = 0; i < nviews; ++i) PyBuffer_Release(&views[i]);
    free(views); free(pq); free(pt); free(pc); free(laln); free(rows);
    Py_XDECREF(poff); Py_XDECREF(soff);
    Py_XDECREF(fq); Py_XDECREF(ft); Py_XDECREF(fc);
    return result;
}

static PyMethodDef methods[] = {
    {"cmap_ragged", cmap_ragged, METH_VARARGS, "collect per-alignment pointers and call mdf_cmap_build_transfer_ragged (GIL released)"},
    {"submit_ragged", submit_ragged, METH_VARARGS, "collect per-protein pointers and call mdf_path_submit_ragged (GIL released)"},
    {"pointers", pointers, METH_O, "pointer / length arrays of a list of str or bytes"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_mdf_pyhost", "list -> pointer-array glue for libmdf_b200", -1, methods};

PyMODINIT_FUNC PyInit__mdf_pyhost(void) { return PyModule_Create(&moddef); }
