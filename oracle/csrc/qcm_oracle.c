/*
 * CPU oracle / CPU baseline for the QCMRF statevector path -- plain C + OpenMP.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Never linked into, or
 * called by, the product library.
 *
 * Restates, as one OpenMP sweep over a dense little-endian complex128 state per
 * gate, what the reference gets from qiskit-aer's qasm_simulator
 * (/root/reference/run_experiment.py:54-57) for the program QCMRF._build emits
 * (/root/reference/QCMRF.py:199-243).  qiskit-aer is an un-vendored third-party
 * dependency that is absent here; this is a "port" baseline, not Aer itself.
 *
 * Op kinds
 *   ORC_U1    2x2 complex matrix on `target` where (idx & cmask) == cval
 *             (h, x, sx, rz, cx, and AND/mcx with open/closed controls)
 *   ORC_PHASE multiply by e^{i lam} where (idx & cmask) == cval  (cp, p, cz)
 *   ORC_MUX   uniformly-controlled 2x2 on `target`: matrix = table[t], table index
 *             bit j = bit ctrls[j] of idx  (the fused clique block, B2 baseline)
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ORC_U1 = 0, ORC_PHASE = 1, ORC_MUX = 2 };

typedef struct {
    int32_t kind;
    int32_t target;
    uint64_t cmask;
    uint64_t cval;
    double m[8];        /* U1: row-major 2x2 (re,im) ; PHASE: m[0] = lam */
    int32_t nctrl;
    int32_t ctrls[8];
    int64_t tab_off;    /* MUX: offset (in doubles) of 2^nctrl * 8 doubles */
} orc_op;

static inline uint64_t insert_zero(uint64_t p, int t) {
    uint64_t lo = p & ((1ull << t) - 1);
    return ((p >> t) << (t + 1)) | lo;
}

int orc_sizeof_op(void) { return (int)sizeof(orc_op); }

int orc_init_zero_state(int N, double complex *psi) {
    uint64_t dim = 1ull << N;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < dim; ++i) psi[i] = 0.0;
    psi[0] = 1.0;
    return 0;
}

int orc_run(int N, const orc_op *ops, int n_ops, const double *tables, double complex *psi) {
    const uint64_t dim = 1ull << N, half = dim >> 1;
    for (int g = 0; g < n_ops; ++g) {
        const orc_op *op = &ops[g];
        if (op->kind == ORC_U1) {
            const double complex u00 = op->m[0] + I * op->m[1], u01 = op->m[2] + I * op->m[3];
            const double complex u10 = op->m[4] + I * op->m[5], u11 = op->m[6] + I * op->m[7];
            const int t = op->target;
            const uint64_t cm = op->cmask, cv = op->cval, st = 1ull << t;
#pragma omp parallel for schedule(static)
            for (uint64_t p = 0; p < half; ++p) {
                uint64_t i0 = insert_zero(p, t);
                if ((i0 & cm) != cv) continue;
                double complex a0 = psi[i0], a1 = psi[i0 | st];
                psi[i0] = u00 * a0 + u01 * a1;
                psi[i0 | st] = u10 * a0 + u11 * a1;
            }
        } else if (op->kind == ORC_PHASE) {
            const double complex ph = cos(op->m[0]) + I * sin(op->m[0]);
            const uint64_t cm = op->cmask, cv = op->cval;
#pragma omp parallel for schedule(static)
            for (uint64_t i = 0; i < dim; ++i)
                if ((i & cm) == cv) psi[i] *= ph;
        } else if (op->kind == ORC_MUX) {
            const int t = op->target, nc = op->nctrl;
            const uint64_t st = 1ull << t;
            const double *tab = tables + op->tab_off;
#pragma omp parallel for schedule(static)
            for (uint64_t p = 0; p < half; ++p) {
                uint64_t i0 = insert_zero(p, t);
                unsigned ti = 0;
                for (int j = 0; j < nc; ++j) ti |= (unsigned)((i0 >> op->ctrls[j]) & 1ull) << j;
                const double *m = tab + 8 * (size_t)ti;
                double complex a0 = psi[i0], a1 = psi[i0 | st];
                psi[i0] = (m[0] + I * m[1]) * a0 + (m[2] + I * m[3]) * a1;
                psi[i0 | st] = (m[4] + I * m[5]) * a0 + (m[6] + I * m[7]) * a1;
            }
        } else {
            return -1;
        }
    }
    return 0;
}

/* masked probability: sum |psi|^2 over (idx & mask) == value; optional copy of the
 * first 2^n_out probabilities (the post-selected block, eval.py:116-123). */
int orc_postselect(int N, const double complex *psi, uint64_t mask, uint64_t value,
                   int n_out, double *probs_out, double *kept_out) {
    const uint64_t dim = 1ull << N;
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (uint64_t i = 0; i < dim; ++i)
        if ((i & mask) == value) acc += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
    if (probs_out)
        for (uint64_t i = 0; i < (1ull << n_out); ++i)
            probs_out[i] = creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
    *kept_out = acc;
    return 0;
}

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* multinomial shots by inverse-CDF: chunk sums -> prefix -> search (what Aer's
 * sample_measure does on the final probabilities). indices_out[shots] */
int orc_sample(int N, const double complex *psi, uint64_t shots, uint64_t seed, uint64_t *indices_out) {
    const uint64_t dim = 1ull << N;
    const int cb = N > 12 ? 12 : N;
    const uint64_t csz = 1ull << cb, nch = dim >> cb;
    double *cs = (double *)malloc(sizeof(double) * (nch + 1));
    if (!cs) return -2;
#pragma omp parallel for schedule(static)
    for (uint64_t c = 0; c < nch; ++c) {
        double a = 0.0;
        for (uint64_t i = c * csz; i < (c + 1) * csz; ++i)
            a += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
        cs[c + 1] = a;
    }
    cs[0] = 0.0;
    for (uint64_t c = 0; c < nch; ++c) cs[c + 1] += cs[c];
    const double total = cs[nch];
#pragma omp parallel for schedule(static)
    for (uint64_t s = 0; s < shots; ++s) {
        uint64_t st = seed ^ (0xd1342543de82ef95ull * (s + 1));
        double u = (double)(splitmix64(&st) >> 11) * (1.0 / 9007199254740992.0) * total;
        uint64_t lo = 0, hi = nch;           /* last c with cs[c] <= u */
        while (hi - lo > 1) {
            uint64_t mid = (lo + hi) >> 1;
            if (cs[mid] <= u) lo = mid; else hi = mid;
        }
        double acc = cs[lo];
        uint64_t i = lo * csz, end = (lo + 1) * csz, pick = lo * csz;
        for (; i < end; ++i) {
            double w = creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
            if (w > 0.0) pick = i;           /* never land on a zero-probability state */
            acc += w;
            if (acc > u) break;
        }
        indices_out[s] = pick;
    }
    free(cs);
    return 0;
}
