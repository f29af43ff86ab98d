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

static char bits8[256][8];
static int bits8_ready = 0;
static void init_bits8(void) {
    for (int b = 0; b < 256; ++b)
        for (int j = 0; j < 8; ++j) bits8[b][7 - j] = (char)('0' + ((b >> j) & 1));
    bits8_ready = 1;
}

/* keys[n] (bit c = sampled value of clbit c) -> new dict {width-character bit string: count},
 * keys in ascending numeric order (the order np.unique gives the Python fallback). */
PyObject *qcm_counts_dict(const uint64_t *keys, Py_ssize_t n, int width) {
    if (width < 1) width = 1;
    if (width > 64 || n < 0) {
        PyErr_SetString(PyExc_ValueError, "qcm_counts_dict: width must be 1..64");
        return NULL;
    }
    if (n == 0) return PyDict_New();
    if (!bits8_ready) init_bits8();                       /* called with the GIL held (PyDLL): no race */
    uint64_t *a = (uint64_t *)malloc(2 * (size_t)n * sizeof(uint64_t));
    if (!a) return PyErr_NoMemory();
    memcpy(a, keys, (size_t)n * sizeof(uint64_t));
    radix_sort_u64(a, a + n, (size_t)n, width);
    /* the table is sized once for the number of distinct keys: a dict grown entry by entry is rebuilt ~12 times on the
     * way to 10^4 entries, a third of this function's time */
    Py_ssize_t distinct = 1;
    for (Py_ssize_t k = 1; k < n; ++k) distinct += a[k] != a[k - 1];
    PyObject *d = _PyDict_NewPresized(distinct);
    if (!d) {
        free(a);
        return NULL;
    }
    Py_ssize_t i = 0;
    while (i < n) {
        const uint64_t v = a[i];
        Py_ssize_t j = i + 1;
        while (j < n && a[j] == v) ++j;
        /* the key is ASCII by construction: fill a compact 1-byte string in place (no decoding pass) */
        PyObject *k = PyUnicode_New(width, 127);
        if (k) {
            Py_UCS1 *d8 = PyUnicode_1BYTE_DATA(k);
            /* eight characters per step from a table of the 256 byte patterns (most significant bit first) */
            int c = 0;
            for (; c + 8 <= width; c += 8) memcpy(d8 + width - 8 - c, bits8[(v >> c) & 0xffu], 8);
            for (; c < width; ++c) d8[width - 1 - c] = (Py_UCS1)('0' + ((v >> c) & 1u));
        }
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
 * the first gate that was NOT applied (a gate touching a qubit outside the block, or `limit`).
 *
 * Nine gates in ten of a transpiled circuit are rz (a phase per row) or x / cx (a row permutation): those are kept
 * LAZY -- logical row r is  e^{i phase[r]} * U[perm[r]]  -- and cost O(rows) instead of O(rows x cols); only sx mixes
 * rows (it first folds the two rows' pending phases into its coefficients).  The lazy state is folded back into U
 * before returning. */
int64_t qcm_basis_apply_run(double *U, int64_t n_rows, int64_t n_cols, const int8_t *kind, const int32_t *tq,
                            const int32_t *cq, const double *par, int64_t j, int64_t limit, const int32_t *pos) {
    if (n_rows > 65536 || n_rows < 1) return -1;
    int32_t *perm = (int32_t *)malloc((size_t)n_rows * sizeof(int32_t));
    double *phase = (double *)malloc((size_t)n_rows * sizeof(double));
    if (!perm || !phase) {
        free(perm);
        free(phase);
        return -1;
    }
    for (int64_t r = 0; r < n_rows; ++r) {
        perm[r] = (int32_t)r;
        phase[r] = 0.0;
    }
    int dirty = 0;
    int64_t rc = 0;
    for (; j < limit; ++j) {
        const int pt = pos[tq[j]];
        if (pt < 0) break;
        const int64_t tb = (int64_t)1 << pt;
        const int k = kind[j];
        if (k == 3) continue;
        if (k == 0) {                                     /* rz(l) = diag(e^{-il/2}, e^{il/2}) */
            const double h = 0.5 * par[j];
            for (int64_t r = 0; r < n_rows; ++r) phase[r] += (r & tb) ? h : -h;
            dirty = 1;
        } else if (k == 2 || k == 4) {                    /* x / cx: swap the logical rows of every selected pair */
            int64_t cb = 0;
            if (k == 4) {
                const int pc = pos[cq[j]];
                if (pc < 0) break;
                cb = (int64_t)1 << pc;
            }
            for (int64_t r = 0; r < n_rows; ++r) {
                if ((r & tb) || (r & cb) != cb) continue;
                const int32_t tp = perm[r];
                perm[r] = perm[r | tb];
                perm[r | tb] = tp;
                const double tf = phase[r];
                phase[r] = phase[r | tb];
                phase[r | tb] = tf;
            }
            dirty = 1;
        } else if (k == 1) {                              /* sx = 0.5 [[1+i, 1-i], [1-i, 1+i]] on rows e^{i f0} u0, e^{i f1} u1 */
            for (int64_t r = 0; r < n_rows; ++r) {
                if (r & tb) continue;
                const double c0 = cos(phase[r]), s0 = sin(phase[r]), c1 = cos(phase[r | tb]), s1 = sin(phase[r | tb]);
                /* a = (1+i)/2 e^{i f0}, b = (1-i)/2 e^{i f1}, c = (1-i)/2 e^{i f0}, d = (1+i)/2 e^{i f1} */
                const double ar = 0.5 * (c0 - s0), ai = 0.5 * (c0 + s0), br = 0.5 * (c1 + s1), bi = 0.5 * (s1 - c1);
                const double cr = 0.5 * (c0 + s0), ci = 0.5 * (s0 - c0), dr = 0.5 * (c1 - s1), di = 0.5 * (c1 + s1);
                double *p0 = U + 2 * (size_t)perm[r] * (size_t)n_cols;
                double *p1 = U + 2 * (size_t)perm[r | tb] * (size_t)n_cols;
                for (int64_t c = 0; c < n_cols; ++c) {
                    const double x0 = p0[2 * c], y0 = p0[2 * c + 1], x1 = p1[2 * c], y1 = p1[2 * c + 1];
                    p0[2 * c] = (ar * x0 - ai * y0) + (br * x1 - bi * y1);
                    p0[2 * c + 1] = (ar * y0 + ai * x0) + (br * y1 + bi * x1);
                    p1[2 * c] = (cr * x0 - ci * y0) + (dr * x1 - di * y1);
                    p1[2 * c + 1] = (cr * y0 + ci * x0) + (dr * y1 + di * x1);
                }
                phase[r] = 0.0;
                phase[r | tb] = 0.0;
            }
        } else {
            rc = -1;
            break;
        }
    }
    if (dirty && rc == 0) {                               /* fold the pending phases and the row permutation back into U */
        double *tmp = (double *)malloc(2 * (size_t)n_rows * (size_t)n_cols * sizeof(double));
        if (!tmp) {
            rc = -1;
        } else {
            for (int64_t r = 0; r < n_rows; ++r) {
                const double c = cos(phase[r]), sn = sin(phase[r]);
                const double *src = U + 2 * (size_t)perm[r] * (size_t)n_cols;
                double *dst = tmp + 2 * (size_t)r * (size_t)n_cols;
                if (phase[r] == 0.0) {
                    memcpy(dst, src, 2 * (size_t)n_cols * sizeof(double));
                } else {
                    for (int64_t cc = 0; cc < n_cols; ++cc) {
                        const double x = src[2 * cc], y = src[2 * cc + 1];
                        dst[2 * cc] = x * c - y * sn;
                        dst[2 * cc + 1] = x * sn + y * c;
                    }
                }
            }
            memcpy(U, tmp, 2 * (size_t)n_rows * (size_t)n_cols * sizeof(double));
            free(tmp);
        }
    }
    free(perm);
    free(phase);
    return rc < 0 ? rc : j;
}
