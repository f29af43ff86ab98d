/* Host-side result formatting for qcmrf_b200 (CPython C API, loaded with ctypes.PyDLL).
 *
 * Result.get_counts() of the reference stack returns {bitstring: count} with clbit width-1 leftmost
 * (run_experiment.py:57).  Building that dict from the sampled keys is the one
 * per-shot piece of host work on the path; in Python it costs a slice + a dict insert per distinct
 * key (~3 ms for 10^4 shots of a 34-clbit circuit), here it is a radix sort, a run-length pass and
 * one PyDict_SetItem per distinct key.  Not on the GPU path; no CUDA in this file. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void radix_sort_u64(uint64_t *a, uint64_t *tmp, size_t n, int bits) {
    for (int shift = 0; shift < bits; shift += 8) {
        size_t count[257];
        memset(count, 0, sizeof count);
        for (size_t i = 0; i < n; ++i) count[((a[i] >> shift) & 0xff) + 1]++;
        for (int b = 0; b < 256; ++b) count[b + 1] += count[b];
        for (size_t i = 0; i < n; ++i) tmp[count[(a[i] >> shift) & 0xff]++] = a[i];
        memcpy(a, tmp, n * sizeof(uint64_t));
    }
}

/* keys[n] (bit c = sampled value of clbit c) -> new dict {width-character bit string: count},
 * keys in ascending numeric order (the order np.unique gives the Python fallback). */
PyObject *qcm_counts_dict(const uint64_t *keys, Py_ssize_t n, int width) {
    if (width < 1) width = 1;
    if (width > 64 || n < 0) {
        PyErr_SetString(PyExc_ValueError, "qcm_counts_dict: width must be 1..64");
        return NULL;
    }
    PyObject *d = PyDict_New();
    if (!d || n == 0) return d;
    uint64_t *a = (uint64_t *)malloc(2 * (size_t)n * sizeof(uint64_t));
    if (!a) {
        Py_DECREF(d);
        return PyErr_NoMemory();
    }
    memcpy(a, keys, (size_t)n * sizeof(uint64_t));
    radix_sort_u64(a, a + n, (size_t)n, width);
    char buf[65];
    Py_ssize_t i = 0;
    while (i < n) {
        const uint64_t v = a[i];
        Py_ssize_t j = i + 1;
        while (j < n && a[j] == v) ++j;
        for (int c = 0; c < width; ++c) buf[width - 1 - c] = (char)('0' + ((v >> c) & 1u));
        PyObject *k = PyUnicode_FromStringAndSize(buf, width);
        PyObject *cnt = PyLong_FromSsize_t(j - i);
        if (!k || !cnt || PyDict_SetItem(d, k, cnt) < 0) {
            Py_XDECREF(k);
            Py_XDECREF(cnt);
            Py_DECREF(d);
            free(a);
            return NULL;
        }
        Py_DECREF(k);
        Py_DECREF(cnt);
        i = j;
    }
    free(a);
    return d;
}

/* Measure-and-release width: full-width keys of a circuit whose released qubits are
 * not stored.  raw[s] = sampled basis state of the stored qubits (from the GPU sampler);
 * released qubit k reads 1 when u[k][s] < p1_k[index bits of raw[s] at ctrl_k]; stored qubits are copied
 * from raw.  One pass over the shots instead of ~10 numpy temporaries per qubit.
 * ctrl: [nv][max_ctrl] physical positions (first n_ctrl[k] valid); p1: concatenated tables, p1_off[k]
 * their starts; vclbit[k]: clbit of released qubit k or -1; clbit_pos[c]: physical position feeding
 * clbit c or -1 (released / unmeasured / never stored).                                              */
int qcm_released_keys(const int64_t *raw, int64_t shots, const double *u, int nv, const int32_t *n_ctrl,
                      const int32_t *ctrl, int max_ctrl, const double *p1, const int64_t *p1_off,
                      const int32_t *vclbit, const int32_t *clbit_pos, int n_clbits, uint64_t *keys_out) {
    if (!raw || !keys_out || shots < 0 || nv < 0 || n_clbits < 0 || n_clbits > 64) return -1;
    for (int64_t s = 0; s < shots; ++s) {
        const uint64_t r = (uint64_t)raw[s];
        uint64_t key = 0;
        for (int c = 0; c < n_clbits; ++c)
            if (clbit_pos[c] >= 0) key |= ((r >> clbit_pos[c]) & 1ull) << c;
        for (int k = 0; k < nv; ++k) {
            if (vclbit[k] < 0) continue;
            uint32_t idx = 0;
            const int32_t *ck = ctrl + (size_t)k * max_ctrl;
            for (int j = 0; j < n_ctrl[k]; ++j) idx |= (uint32_t)((r >> ck[j]) & 1ull) << j;
            if (u[(size_t)k * shots + s] < p1[p1_off[k] + idx]) key |= 1ull << vclbit[k];
        }
        keys_out[s] = key;
    }
    return 0;
}

/* Gate-fusion pass (qcmrf_b200/fusion.py, _Block.apply): apply one (multi-)controlled single-qubit gate
 * to the rows of the block matrix U (complex128, row-major, n_rows x n_cols): for every row r with the
 * target bit clear and (r & cmask) == cval, mix rows r and r | tbit with the 2x2 matrix b (row-major
 * re, im pairs).  The same arithmetic as the numpy version, one pass, no temporaries: a transpiled
 * fixture circuit is ~15 000 of these.                                                              */
void qcm_block_apply(double *U, int64_t n_rows, int64_t n_cols, uint64_t cmask, uint64_t cval, uint64_t tbit,
                     const double *b) {
    const double b00r = b[0], b00i = b[1], b01r = b[2], b01i = b[3], b10r = b[4], b10i = b[5], b11r = b[6], b11i = b[7];
    const int diag = (b01r == 0.0 && b01i == 0.0 && b10r == 0.0 && b10i == 0.0);
    const int xtype = (b00r == 0.0 && b00i == 0.0 && b11r == 0.0 && b11i == 0.0 && b01r == 1.0 && b01i == 0.0 &&
                       b10r == 1.0 && b10i == 0.0);
    for (int64_t r = 0; r < n_rows; ++r) {
        if (((uint64_t)r & tbit) || (((uint64_t)r & cmask) != cval)) continue;
        double *p0 = U + 2 * (size_t)r * (size_t)n_cols;
        double *p1 = U + 2 * (size_t)((uint64_t)r | tbit) * (size_t)n_cols;
        if (diag) {
            const int s0 = !(b00r == 1.0 && b00i == 0.0), s1 = !(b11r == 1.0 && b11i == 0.0);
            for (int64_t c = 0; c < n_cols; ++c) {
                if (s0) {
                    const double x = p0[2 * c], y = p0[2 * c + 1];
                    p0[2 * c] = x * b00r - y * b00i;
                    p0[2 * c + 1] = x * b00i + y * b00r;
                }
                if (s1) {
                    const double x = p1[2 * c], y = p1[2 * c + 1];
                    p1[2 * c] = x * b11r - y * b11i;
                    p1[2 * c + 1] = x * b11i + y * b11r;
                }
            }
        } else if (xtype) {
            for (int64_t c = 0; c < 2 * n_cols; ++c) {
                const double t = p0[c];
                p0[c] = p1[c];
                p1[c] = t;
            }
        } else {
            for (int64_t c = 0; c < n_cols; ++c) {
                const double x0 = p0[2 * c], y0 = p0[2 * c + 1], x1 = p1[2 * c], y1 = p1[2 * c + 1];
                p0[2 * c] = (b00r * x0 - b00i * y0) + (b01r * x1 - b01i * y1);
                p0[2 * c + 1] = (b00r * y0 + b00i * x0) + (b01r * y1 + b01i * x1);
                p1[2 * c] = (b10r * x0 - b10i * y0) + (b11r * x1 - b11i * y1);
                p1[2 * c + 1] = (b10r * y0 + b10i * x0) + (b11r * y1 + b11i * x1);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------
 * Basis-gate programs (the output of transpile(..., basis_gates=['cx','id','rz','sx','x']),
 * /root/reference/run_experiment.py:52) as flat arrays: kind[] (0 rz, 1 sx, 2 x, 3 id, 4 cx), tq[] target,
 * cq[] control (cx) or -1, par[] angle (rz).  A transpiled fixture circuit has up to ~15 000 gates; walking them
 * one Python object at a time cost ~110 ms per circuit in the gate-fusion pass -- these two helpers keep the
 * per-gate work in C and leave the Python side one step per QUBIT that joins a block.
 * ------------------------------------------------------------------------------------------------------ */

/* keep[g] = 0 for a cx whose control no kept gate has targeted yet (the qubit is still |0>: a closed control on it
 * never fires), else 1 -- the array twin of fusion.prune_zero_controls.  Returns the number of kept gates. */
int64_t qcm_basis_prune(const int8_t *kind, const int32_t *tq, const int32_t *cq, int64_t n, int32_t n_qubits, uint8_t *keep) {
    uint8_t *clean = (uint8_t *)malloc((size_t)(n_qubits > 0 ? n_qubits : 1));
    if (!clean) return -1;
    memset(clean, 1, (size_t)(n_qubits > 0 ? n_qubits : 1));
    int64_t kept = 0;
    for (int64_t g = 0; g < n; ++g) {
        if (kind[g] == 4 && cq[g] >= 0 && cq[g] < n_qubits && clean[cq[g]]) {
            keep[g] = 0;
            continue;
        }
        keep[g] = 1;
        ++kept;
        if (tq[g] >= 0 && tq[g] < n_qubits) clean[tq[g]] = 0;
    }
    free(clean);
    return kept;
}

/* Applies gates j, j+1, ... (< limit) to the block matrix U (complex128, row-major, n_rows x n_cols; row bit p <->
 * block position p) for as long as every qubit of the gate has a block position (pos[q] >= 0); returns the index of
 * the first gate that was NOT applied (a gate touching a qubit outside the block, or `limit`). */
int64_t qcm_basis_apply_run(double *U, int64_t n_rows, int64_t n_cols, const int8_t *kind, const int32_t *tq,
                            const int32_t *cq, const double *par, int64_t j, int64_t limit, const int32_t *pos) {
    static const double X[8] = {0, 0, 1, 0, 1, 0, 0, 0};
    static const double SX[8] = {0.5, 0.5, 0.5, -0.5, 0.5, -0.5, 0.5, 0.5};
    for (; j < limit; ++j) {
        const int pt = pos[tq[j]];
        if (pt < 0) break;
        switch (kind[j]) {
            case 0: {                                     /* rz(l) = diag(e^{-il/2}, e^{il/2}) */
                const double h = 0.5 * par[j];
                const double b[8] = {cos(h), -sin(h), 0, 0, 0, 0, cos(h), sin(h)};
                qcm_block_apply(U, n_rows, n_cols, 0, 0, 1ull << pt, b);
                break;
            }
            case 1: qcm_block_apply(U, n_rows, n_cols, 0, 0, 1ull << pt, SX); break;
            case 2: qcm_block_apply(U, n_rows, n_cols, 0, 0, 1ull << pt, X); break;
            case 3: break;
            case 4: {
                const int pc = pos[cq[j]];
                if (pc < 0) return j;
                qcm_block_apply(U, n_rows, n_cols, 1ull << pc, 1ull << pc, 1ull << pt, X);
                break;
            }
            default: return -1;
        }
    }
    return j;
}
