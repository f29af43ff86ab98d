"""Gate fusion and pass planning (host side of the hot path).

Stage 1 ``fuse``: greedy qubit-set block fusion with structure classification.
  Consecutive gates are accumulated into a small dense unitary over the qubits they
  touch (<= ``q_max``); at every point where the next gate would bring in a new
  qubit the block is *classified*, restricted to what is known about its inputs
  (qubits never touched so far are |0>):
    - qubits that provably return to |0> are dropped from the block (the AND
      scratch qubit n of QCMRF.py:219-227 disappears here),
    - what is left must be a diagonal (DIAG) or a uniformly-controlled single-qubit
      gate (MUX1Q: block-diagonal over every qubit but one).
  The longest prefix that classifies is emitted as ONE sweep.  For the program
  QCMRF._build emits (QCMRF.py:216-236) -- and equally for its ``transpile``d
  cx/rz/sx/x form -- every clique block becomes one MUX1Q on the clique's ancilla,
  controlled by the clique's variable qubits: the uniformly-controlled RX(4 gamma)
  of SURVEY.md App. A, found numerically rather than assumed.
  Single-qubit gates on still-|0> qubits fold into the product-state initialiser.

Stage 2 ``plan``: physical layout + lazy materialisation + blocked passes.
  Qubits are laid out in order of first use, so the state grows from 2^n_init
  amplitudes and a qubit that is never materialised (the scratch qubit) is never
  stored.  Consecutive MUX1Q sweeps on distinct targets whose index qubits are not
  block targets are merged into BLOCK passes of up to ``block_max`` targets, each
  amplitude crossing HBM once per pass.

Nothing here touches amplitudes: the output is a list of ``qcm_op`` + coefficient
tables for the CUDA engine.
"""
import cmath
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .ir import Gate, Program
from .ir import _CTRL_BASE as _CTRL_BASE_NAME

TOL = 1e-9

QCM_OP_INIT_PRODUCT, QCM_OP_MUX1Q, QCM_OP_DIAG, QCM_OP_BLOCK, QCM_OP_SWAP, QCM_OP_EXTEND = 1, 2, 3, 4, 5, 6
QCM_MAX_CTRL, QCM_MAX_BLOCK, QCM_MAX_MEMBERS, QCM_MAX_EXPAND = 10, 5, 16, 8


@dataclass
class FusedOp:
    kind: str                       # 'mux' | 'diag'
    target: int                     # logical qubit ('mux'), -1 for 'diag'
    ctrls: Tuple[int, ...]          # logical qubits; table-index bit j <-> ctrls[j]
    table: np.ndarray               # mux: (2^m, 2, 2) ; diag: (2^m,)   complex128
    zero_in: bool = False           # mux: target known |0> on input
    n_gates: int = 0                # primitive gates folded into this op


@dataclass
class FusedCircuit:
    n_qubits: int
    init: Dict[int, np.ndarray]     # logical qubit -> its 2-vector after folded 1q gates
    ops: List[FusedOp]
    global_phase: float = 0.0
    n_gates_in: int = 0


# ------------------------------------------------------------------------------------------
class _Block:
    """Action of a run of gates on a growing little-endian qubit list, restricted to
    the inputs that can occur: qubits known to be |0> when they join contribute no
    input columns.  U has one row per basis state of all block qubits and one column
    per basis state of the block qubits that were NOT known-|0>."""

    def __init__(self, zero):
        self.zero = zero
        self.qubits: List[int] = []
        self.pos: Dict[int, int] = {}
        self.cpos: List[int] = []                  # block positions that own a column bit
        self.U = np.ones((1, 1), dtype=np.complex128)
        self.n_gates = 0
        self._uaddr_of, self._uaddr = None, 0      # address of self.U for the C row mixer (refreshed when U is replaced)

    def add_qubit(self, q):
        p = len(self.qubits)
        self.pos[q] = p
        self.qubits.append(q)
        if q in self.zero:
            self.U = np.concatenate([self.U, np.zeros_like(self.U)], axis=0)
        else:
            self.cpos.append(p)
            self.U = np.kron(np.eye(2, dtype=np.complex128), self.U)

    def apply(self, g: Gate):
        for q in g.qubits:
            if q not in self.pos:
                self.add_qubit(q)
        nq = len(self.qubits)
        B = g.base_matrix()
        U = self.U
        cmask = cval = 0
        for q, v in zip(g.controls, g.ctrl_values):
            cmask |= 1 << self.pos[q]
            cval |= v << self.pos[q]
        if _HOST_APPLY is not None:
            # one C pass over the selected row pairs (csrc/qcm_host.c), same arithmetic as below
            if self._uaddr_of is not U:
                if not U.flags.c_contiguous:
                    U = self.U = np.ascontiguousarray(U)
                self._uaddr_of, self._uaddr = U, U.ctypes.data
            bk = (g.name if not g.controls else _CTRL_BASE_NAME.get(g.name, g.name), g.params)
            ent = _B_ADDR.get(bk)
            if ent is None or ent[0] is not B:
                Bc = np.ascontiguousarray(B, dtype=np.complex128)
                ent = _B_ADDR[bk] = (B, Bc, Bc.ctypes.data)
                if len(_B_ADDR) > 65536:
                    _B_ADDR.clear()
            _HOST_APPLY(self._uaddr, U.shape[0], U.shape[1], cmask, cval, 1 << self.pos[g.target], ent[2])
            self.n_gates += 1
            return
        r0, r1 = _pair_rows(nq, cmask, cval, 1 << self.pos[g.target])
        b00, b01, b10, b11 = B[0, 0], B[0, 1], B[1, 0], B[1, 1]
        if b01 == 0 and b10 == 0:                       # diagonal: scale rows
            if b00 != 1:
                U[r0] *= b00
            if b11 != 1:
                U[r1] *= b11
        elif b00 == 0 and b11 == 0 and b01 == 1 and b10 == 1:   # X-type: swap rows
            a0 = U[r0]
            U[r0] = U[r1]
            U[r1] = a0
        else:
            a0 = U[r0]
            a1 = U[r1]
            U[r0] = b00 * a0 + b01 * a1
            U[r1] = b10 * a0 + b11 * a1
        self.n_gates += 1


def _load_host_apply():
    """qcm_block_apply of qcmrf_b200/_qcm_host.so, or None when the library is not built (numpy path)."""
    import ctypes
    import os
    try:
        L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '_qcm_host.so'))
        f = L.qcm_block_apply
    except (OSError, AttributeError):
        return None
    f.restype = None
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                  ctypes.c_void_p]
    return f


def _load_host_basis():
    """(qcm_basis_prune, qcm_basis_apply_run) of qcmrf_b200/_qcm_host.so, or None when the library is not built."""
    import ctypes
    import os
    try:
        L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '_qcm_host.so'))
        pr, ar = L.qcm_basis_prune, L.qcm_basis_apply_run
    except (OSError, AttributeError):
        return None
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    pr.restype = i64
    pr.argtypes = [vp, vp, vp, i64, ctypes.c_int32, vp]
    ar.restype = i64
    ar.argtypes = [vp, i64, i64, vp, vp, vp, vp, i64, i64, vp]
    return pr, ar


_HOST_APPLY = _load_host_apply()
_HOST_BASIS = _load_host_basis()
_B_ADDR = {}                                       # (base gate, params) -> (matrix, contiguous copy, its address)
_PAIR_ROWS = {}


def _pair_rows(nq, cmask, cval, tb):
    """Row index pairs (target bit 0 / 1) of a 2^nq-row matrix selected by a control pattern."""
    key = (nq, cmask, cval, tb)
    hit = _PAIR_ROWS.get(key)
    if hit is None:
        rows = np.arange(1 << nq)
        r0 = rows[((rows & cmask) == cval) & ((rows & tb) == 0)]
        hit = (r0, r0 | tb)
        if len(_PAIR_ROWS) < 4096:
            _PAIR_ROWS[key] = hit
    return hit


def _spread(values, positions):
    """Scatter bit j of every value to bit positions[j]."""
    out = np.zeros_like(values)
    for j, p in enumerate(positions):
        out |= ((values >> j) & 1) << p
    return out


def _classify(blk: _Block, tol=TOL):
    """None if the block is not a single sweep, else (ops, touched_qubits, phase).
    Known-|0> qubits that provably return to |0> are dropped from the sweep."""
    Q = blk.qubits
    nq = len(Q)
    U = blk.U
    ridx = np.arange(1 << nq)
    zpos = [p for p in range(nq) if p not in blk.cpos]
    restored = [p for p in zpos if np.abs(U[((ridx >> p) & 1) == 1]).max(initial=0.0) < tol]
    rmask = sum(1 << p for p in restored)
    kpos = [p for p in range(nq) if p not in restored]             # kept block positions
    kof = {p: j for j, p in enumerate(kpos)}
    nk = len(kpos)
    Vc = U[(ridx & rmask) == 0]                                     # rows over kept qubits
    kidx = np.arange(1 << nk)
    ccols = np.arange(U.shape[1])
    cols_k = _spread(ccols, [kof[p] for p in blk.cpos])             # column -> kept-space index
    zk = [kof[p] for p in zpos if p not in restored]                # materialised |0> qubits
    kq = [Q[p] for p in kpos]
    D = kidx[:, None] ^ cols_k[None, :]

    if not zk and np.abs(Vc[D != 0]).max(initial=0.0) < tol:        # ---- diagonal
        d = Vc[cols_k, ccols]
        if np.abs(d - d[0]).max(initial=0.0) < tol:
            return [], [], float(np.angle(d[0]))
        dep = [j for j in range(nk)
               if np.abs(d[((cols_k >> j) & 1) == 0] - d[((cols_k >> j) & 1) == 1]).max() >= tol]
        if len(dep) > QCM_MAX_CTRL:
            return None
        dfull = np.empty(1 << nk, dtype=np.complex128)
        dfull[cols_k] = d
        table = dfull[_spread(np.arange(1 << len(dep)), dep)]
        qs = tuple(kq[j] for j in dep)
        return [FusedOp('diag', -1, qs, table, False, blk.n_gates)], list(qs), 0.0

    col_of = np.full(1 << nk, -1, dtype=np.int64)
    col_of[cols_k] = ccols
    for tp in range(nk):                                            # ---- multiplexer on tp
        if zk and zk != [tp]:
            continue                    # a materialised |0> qubit must be THE target
        tb = 1 << tp
        if np.abs(Vc[(D != 0) & (D != tb)]).max(initial=0.0) >= tol:
            continue
        others = [j for j in range(nk) if j != tp]
        m = len(others)
        i0 = _spread(np.arange(1 << m), others)
        i1 = i0 | tb
        c0 = col_of[i0]
        table = np.zeros((1 << m, 2, 2), dtype=np.complex128)
        table[:, 0, 0] = Vc[i0, c0]
        table[:, 1, 0] = Vc[i1, c0]
        zero_in = tp in zk
        if zero_in:                     # |1> input never occurs: complete the matrix unitarily
            table[:, 0, 1] = -np.conj(table[:, 1, 0])
            table[:, 1, 1] = np.conj(table[:, 0, 0])
        else:
            c1 = col_of[i1]
            table[:, 0, 1] = Vc[i0, c1]
            table[:, 1, 1] = Vc[i1, c1]
        tsel = np.arange(1 << m)
        dep = [jj for jj in range(m)
               if np.abs(table[((tsel >> jj) & 1) == 0] - table[((tsel >> jj) & 1) == 1]).max() >= tol]
        if len(dep) > QCM_MAX_CTRL:
            return None
        table = table[_spread(np.arange(1 << len(dep)), dep)]
        ctrls = tuple(kq[others[jj]] for jj in dep)
        op = FusedOp('mux', kq[tp], ctrls, table, zero_in, blk.n_gates)
        return [op], [kq[tp]] + list(ctrls), 0.0
    return None


def _unrestored_zero_qubits(blk: _Block, tol=TOL):
    """Known-|0> qubits of the block that its gates leave with weight on |1> (for some input).  Two or
    more of them and the block cannot be a single sweep (_classify: such a qubit must be THE target)."""
    nq = len(blk.qubits)
    U = blk.U
    ridx = np.arange(1 << nq)
    rowmax = np.abs(U).max(axis=1, initial=0.0)
    return {blk.qubits[p] for p in range(nq)
            if p not in blk.cpos and rowmax[((ridx >> p) & 1) == 1].max(initial=0.0) >= tol}


_MATCH = {}


def _match_patterns(m, pos, vals):
    """Indices c in [0, 2^m) whose bits at ``pos`` equal ``vals``."""
    key = (m, pos, vals)
    hit = _MATCH.get(key)
    if hit is None:
        c = np.arange(1 << m)
        sel = np.ones(1 << m, dtype=bool)
        for p, v in zip(pos, vals):
            sel &= ((c >> p) & 1) == v
        hit = c[sel]
        if len(_MATCH) < 4096:
            _MATCH[key] = hit
    return hit


_B4_CONST = {}


def _b4(g: Gate):
    """Base 2x2 of a gate as four Python complex numbers (no numpy on the per-gate path)."""
    from .ir import _CTRL_BASE, _ONEQ
    name = _CTRL_BASE.get(g.name, g.name) if len(g.qubits) > 1 else g.name
    if name == 'p':
        return 1 + 0j, 0j, 0j, cmath.exp(1j * g.params[0])
    if name == 'rz':
        return cmath.exp(-0.5j * g.params[0]), 0j, 0j, cmath.exp(0.5j * g.params[0])
    hit = _B4_CONST.get(name)
    if hit is None:
        if name not in _ONEQ:
            B = g.base_matrix()
            return complex(B[0, 0]), complex(B[0, 1]), complex(B[1, 0]), complex(B[1, 1])
        B = _ONEQ[name]
        hit = _B4_CONST[name] = (complex(B[0, 0]), complex(B[0, 1]), complex(B[1, 0]), complex(B[1, 1]))
    return hit


def _run_mux(run: List[Gate], zero_in: bool, tol=TOL):
    """Consecutive gates with one common target: the product is a uniformly-controlled gate on that
    target over the union of their controls.  Built entry by entry from 2x2 products (a gate with a
    full control pattern touches ONE table entry).  None if the union exceeds QCM_MAX_CTRL."""
    t = run[0].qubits[-1]
    ctrls: List[int] = []
    for g in run:
        for q in g.qubits[:-1]:
            if q not in ctrls:
                ctrls.append(q)
    m = len(ctrls)
    if m > QCM_MAX_CTRL:
        return None
    cpos = {q: k for k, q in enumerate(ctrls)}
    tab = [[1 + 0j, 0j, 0j, 1 + 0j] for _ in range(1 << m)]        # row-major 2x2 per control pattern
    for g in run:
        b00, b01, b10, b11 = _b4(g)
        qs = g.qubits
        if len(qs) - 1 == m:
            k = 0
            for q, v in zip(qs, g.ctrl_values):
                k |= v << cpos[q]
            idx = (k,)
        else:
            idx = _match_patterns(m, tuple(cpos[q] for q in qs[:-1]), tuple(g.ctrl_values))
        for k in idx:
            e = tab[k]
            e[0], e[1], e[2], e[3] = (b00 * e[0] + b01 * e[2], b00 * e[1] + b01 * e[3],
                                      b10 * e[0] + b11 * e[2], b10 * e[1] + b11 * e[3])
    # index qubits the table does not depend on (a |0> input only ever sees column 0), in plain Python:
    # the tables have at most a few dozen entries
    cols = (0, 2) if zero_in else (0, 1, 2, 3)
    dep = []
    for j in range(m):
        bit = 1 << j
        if any(abs(tab[c][k] - tab[c | bit][k]) >= tol for c in range(1 << m) if not c & bit for k in cols):
            dep.append(j)
    if len(dep) < m:
        sel = []
        for c in range(1 << len(dep)):
            full = 0
            for jj, j in enumerate(dep):
                full |= ((c >> jj) & 1) << j
            sel.append(tab[full])
        tab = sel
        ctrls = [ctrls[j] for j in dep]
    if not zero_in and not ctrls:
        e = tab[0]
        if abs(e[1]) < tol and abs(e[2]) < tol and abs(e[3] - e[0]) < tol:
            return [], [], float(cmath.phase(e[0]))         # identity up to a phase
    if zero_in:                                            # same completion as _classify: unitary 2x2
        tab = [[e[0], -e[2].conjugate(), e[2], e[0].conjugate()] for e in tab]
    table = np.array(tab, dtype=np.complex128).reshape(len(tab), 2, 2)
    op = FusedOp('mux', t, tuple(ctrls), table, zero_in, len(run))
    return [op], [t] + list(ctrls), 0.0


def _single_gate_op(g: Gate, zero: set):
    blk = _Block(zero)
    blk.apply(g)
    res = _classify(blk)
    if res is None:
        blk = _Block(set())
        blk.apply(g)
        res = _classify(blk)
    assert res is not None, 'primitive gate must classify'
    return res


def direct_ops(prog: Program) -> 'FusedCircuit':
    """One sweep per primitive gate, no matrix classification (fusion='off', and the
    batched small-circuit path where a sweep over shared memory is almost free)."""
    ops: List[FusedOp] = []
    for g in prog.gates:
        B = g.base_matrix()
        m = len(g.controls)
        pat = sum(v << j for j, v in enumerate(g.ctrl_values))
        if B[0, 1] == 0 and B[1, 0] == 0:
            if m == 0 and B[0, 0] == B[1, 1]:
                continue                                           # identity up to phase
            table = np.ones(2 << m, dtype=np.complex128)
            table[pat] = B[0, 0]
            table[pat | (1 << m)] = B[1, 1]
            ops.append(FusedOp('diag', -1, tuple(g.controls) + (g.target,), table, False, 1))
        else:
            table = np.tile(np.eye(2, dtype=np.complex128), (1 << m, 1, 1))
            table[pat] = B
            ops.append(FusedOp('mux', g.target, tuple(g.controls), table, False, 1))
    return FusedCircuit(prog.n_qubits, {}, ops, prog.global_phase, len(prog.gates))


def fold_clean_scratch(gates: List[Gate], n_qubits: int) -> List[Gate]:
    """Compute / use / uncompute on a clean scratch qubit:

        mcx(A == v -> s) ; G_1 .. G_m (s only as a closed control) ; mcx(A == v -> s),  s known |0>

    equals G_1 .. G_m with the control on s replaced by controls A == v, and s is never
    touched.  This is the AND . CP . AND of QCMRF.py:224-227 (one multi-controlled phase per
    clique state); folding it first shrinks every clique block from |C|+2 to |C|+1 qubits and
    a third of the gates before the matrix-based fusion sees it.  Anything else passes through."""
    clean = set(range(n_qubits))
    out: List[Gate] = []
    i, n = 0, len(gates)
    while i < n:
        g = gates[i]
        gq = g.qubits
        s_q = gq[-1]
        folded = False
        if len(gq) > 1 and g.name in ('cx', 'mcx') and s_q in clean:
            A = set(gq[:-1])
            j = i + 1
            if i + 2 < n and gates[i + 2] is g:          # the common shape: compute, ONE use, uncompute
                hq = gates[j].qubits
                if (hq[-1] != s_q and hq[-1] not in A and
                        (s_q not in hq or gates[j].ctrl_values[hq.index(s_q)] == 1)):
                    j = i + 2
            while j < n and j != i + 2:
                hq = gates[j].qubits
                ht = hq[-1]
                if ht == s_q or ht in A:
                    break
                if s_q in hq and gates[j].ctrl_values[hq.index(s_q)] != 1:
                    break
                j += 1
            if j == i + 2 and (j >= n or gates[j].qubits[-1] != s_q):
                j = n                                    # the scan stopped for another reason (or ran off the end)
            if (j < n and j > i + 1 and gates[j].qubits == gq and gates[j].name in ('cx', 'mcx')
                    and gates[j].ctrl_values == g.ctrl_values):
                inner = gates[i + 1:j]
                if all(A.isdisjoint(h.qubits) for h in inner):
                    for h in inner:
                        hq = h.qubits
                        if s_q not in hq:
                            out.append(h)
                            clean.discard(hq[-1])
                            continue
                        k = hq.index(s_q)
                        ctrls = hq[:k] + hq[k + 1:-1] + gq[:-1]
                        vals = h.ctrl_values[:k] + h.ctrl_values[k + 1:] + g.ctrl_values
                        out.append(Gate(ir_ctrl_base(h.name), ctrls + (hq[-1],), h.params, vals))
                        clean.discard(hq[-1])
                    i = j + 1
                    folded = True
        if not folded:
            out.append(g)
            clean.discard(s_q)
            i += 1
    return out


def prune_zero_controls(gates: List[Gate], n_qubits: int) -> List[Gate]:
    """A qubit no gate has targeted yet is still |0>: a closed control on it never fires (the gate is
    the identity and is dropped), an open control on it always fires (the control is dropped).  Without
    this a sweep could be indexed by a qubit the lazy layout never materialises (`cx(2,1)` on |0000>)."""
    clean = set(range(n_qubits))
    out: List[Gate] = []
    for g in gates:
        gq = g.qubits
        if len(gq) > 1 and not clean.isdisjoint(gq[:-1]):
            if any(q in clean and v == 1 for q, v in zip(gq, g.ctrl_values)):
                continue
            keep = [(q, v) for q, v in zip(gq, g.ctrl_values) if q not in clean]
            if keep:
                g = Gate(g.name, tuple(q for q, _ in keep) + (gq[-1],), g.params, tuple(v for _, v in keep))
            else:
                g = Gate(_CTRL_BASE_NAME.get(g.name, g.name), (gq[-1],), g.params, ())
        out.append(g)
        clean.discard(gq[-1])
    return out


def drop_unmaterialised_controls(fc: 'FusedCircuit') -> 'FusedCircuit':
    """A sweep's index qubit that is neither initialised nor the target of an earlier sweep is still |0> when the sweep
    runs: the table is sliced at that index bit = 0 and the qubit leaves the index (a sweep that becomes the identity on a
    |0> target, or a diagonal without index qubits, disappears).  prune_zero_controls does this gate by gate, but a qubit
    whose gates only CLASSIFY to "still |0>" (`cx` from a |0> control, `p`, a controlled phase: found by the seeded fuzz in
    tests/test_host_fusion.py) looked touched to it, and the lazy layout then reserved a position for a qubit that is
    never materialised ('layout/materialisation order mismatch')."""
    mat = set(fc.init)
    ops: List[FusedOp] = []
    phase = fc.global_phase
    changed = False
    for op in fc.ops:
        dead = [j for j, q in enumerate(op.ctrls) if q not in mat]
        if dead:
            changed = True
            keep = [j for j in range(len(op.ctrls)) if j not in dead]
            ar = np.arange(1 << len(keep), dtype=np.int64)
            idx = np.zeros(1 << len(keep), dtype=np.int64)
            for k2, j in enumerate(keep):
                idx |= ((ar >> k2) & 1) << j
            tab = np.ascontiguousarray(op.table[idx])
            ctrls = tuple(op.ctrls[j] for j in keep)
            if op.kind == 'diag' and not ctrls:
                phase += float(np.angle(tab[0]))
                continue
            if op.kind == 'mux' and op.target not in mat and np.abs(tab - np.eye(2)).max() == 0.0:
                continue                                    # the identity on a |0> target: the target stays unmaterialised
            op = FusedOp(op.kind, op.target, ctrls, tab, op.zero_in, op.n_gates)
        ops.append(op)
        if op.kind == 'mux':
            mat.add(op.target)
    if not changed:
        return fc
    import dataclasses
    return dataclasses.replace(fc, ops=ops, global_phase=phase)


_MC_NAME = {'cx': 'mcx', 'mcx': 'mcx', 'cp': 'mcp', 'mcp': 'mcp'}


def ir_ctrl_base(name):
    """Name of a controlled primitive once it has more controls ('cp' -> 'mcp', 'cx' -> 'mcx'; the
    other controlled names already stand for any number of controls in this IR)."""
    return _MC_NAME.get(name, name)


def fuse(prog: Program, mode: str = 'clique', q_max: int = 8, use_hint: bool = True) -> FusedCircuit:
    """mode 'off': one sweep per primitive gate; 'clique': block fusion."""
    if mode == 'off':
        return direct_ops(prog)
    hint = getattr(prog, 'fused_hint', None)
    if hint is not None and q_max == 8 and use_hint:
        fc = hint()                                # the producing circuit class knows its own fused form
        if fc is not None:
            return fc
    if use_hint and _HOST_BASIS is not None and hasattr(prog, 'bk'):
        return _fuse_basis(prog, q_max)            # basis-gate arrays (transpile's output): the per-gate work runs in C
    zero = set(range(prog.n_qubits))
    ops: List[FusedOp] = []
    phase = prog.global_phase
    gates: List[Gate] = prune_zero_controls(fold_clean_scratch(prog.gates, prog.n_qubits), prog.n_qubits)
    n = len(gates)
    last_use: Dict[int, int] = {}
    first_use: Dict[int, int] = {}
    last_target: Dict[int, int] = {}
    for gi, g in enumerate(gates):
        for q in g.qubits:
            last_use[q] = gi
            first_use.setdefault(q, gi)
        last_target[g.qubits[-1]] = gi

    def emit(res):
        nonlocal phase
        new_ops, touched_q, ph = res
        phase += ph
        ops.extend(new_ops)
        zero.difference_update(touched_q)

    def retired(res, j):
        """The classified sweep's target has no gate at or after position j: extending
        the block further cannot keep it a single-target sweep."""
        new_ops = res[0]
        return len(new_ops) == 1 and new_ops[0].kind == 'mux' and last_use.get(new_ops[0].target, -1) < j

    def lifetime_fits(t, i):
        """Do the gates from i to the last use of qubit t touch at most q_max qubits?"""
        seen = set()
        for g in gates[i:last_use[t] + 1]:
            seen.update(g.qubits)
            if len(seen) > q_max:
                return False
        return True

    live_from = [0] * (n + 1)                      # distinct qubits used by gates[i:]
    seen_q: set = set()
    for gi in range(n - 1, -1, -1):
        seen_q.update(gates[gi].qubits)
        live_from[gi] = len(seen_q)
    doom: Optional[set] = None

    i = 0
    while i < n:
        g0 = gates[i]
        if (len(g0.qubits) == 1 and g0.target in zero and
                (last_target[g0.target] == i or not lifetime_fits(g0.target, i))):
            # a lone preparation gate on a qubit that lives too long to be the target of
            # one fused block (the H layer, QCMRF.py:204-205): emit now, it folds into INIT
            B = g0.base_matrix()
            emit(([FusedOp('mux', g0.target, (), np.array([B], dtype=np.complex128), True, 1)], [g0.target], 0.0))
            i += 1
            continue
        # a qubit whose whole life is one run of consecutive gates targeting it (a clique ancilla
        # after fold_clean_scratch): the run IS a multiplexer on it -- no dense block needed
        t = g0.target
        if first_use[t] == i:
            j = i
            while j < n and gates[j].target == t:
                j += 1
            if last_use[t] == j - 1:
                res = _run_mux(gates[i:j], t in zero)
                if res is not None:
                    emit(res)
                    i = j
                    continue
        blk = _Block(zero)
        best = None
        j = i
        # `doom`: known-|0> qubits that an earlier block reaching the end of the circuit left on |1>.  If two
        # of them are still known-|0> here, this block's end cannot classify either (they stay unrestored
        # for the larger input set too, and such a qubit must be THE target of a sweep): once every qubit
        # that can still join has joined, applying the tail is pointless.  A transpiled circuit otherwise
        # re-applies its whole body once per variable of the H layer (rz.sx.rz runs peel one per attempt).
        doomed = doom is not None and sum(1 for q in doom if q in zero) >= 2
        while j < n:
            g = gates[j]
            new = [q for q in g.qubits if q not in blk.pos]
            if new and blk.qubits:
                res = _classify(blk)
                if res is not None:
                    best = (j, res)
                    if retired(res, j):
                        break
                if len(blk.qubits) + len(new) > q_max:
                    break
            blk.apply(g)
            j += 1
            if doomed and len(blk.qubits) == live_from[i]:
                break
        else:
            res = _classify(blk)
            if res is not None:
                best = (j, res)
            else:
                left = _unrestored_zero_qubits(blk)
                doom = left if len(left) >= 2 else None
        if best is None:
            best = (i + 1, _single_gate_op(gates[i], zero))
        emit(best[1])
        i = best[0]

    return _fold_init(prog.n_qubits, ops, phase, len(prog.gates))


def _fuse_basis(prog, q_max: int = 8) -> FusedCircuit:
    """``fuse(prog, 'clique')`` for a basis-gate program held as flat arrays (ir.BasisProgram: the output of
    ``transpile``, run_experiment.py:52).  Same greedy algorithm, same classification (_classify), same result --
    tests/test_host_fusion.py pins it against the object-per-gate loop above -- but the per-gate work (zero-control
    pruning, applying a run of gates to the block matrix) happens in C over the arrays; Python advances one step per
    qubit that joins a block.  A transpiled fixture circuit (up to ~15 000 gates) costs a few ms instead of ~100."""
    prune, apply_run = _HOST_BASIS
    n0 = len(prog.bk)
    keep = np.ones(n0, dtype=np.uint8)
    if n0:
        prune(prog.bk.ctypes.data, prog.bq.ctypes.data, prog.bc.ctypes.data, n0, prog.n_qubits, keep.ctypes.data)
    sel = keep.astype(bool)
    bk = np.ascontiguousarray(prog.bk[sel])
    bq = np.ascontiguousarray(prog.bq[sel])
    bc = np.ascontiguousarray(prog.bc[sel])
    bp = np.ascontiguousarray(prog.bp[sel])
    n = len(bk)
    N = prog.n_qubits
    zero = set(range(N))
    ops: List[FusedOp] = []
    phase = prog.global_phase
    idx = np.arange(n)
    # last / first use of every qubit (as target or control), last gate targeting it
    last_use = np.full(N, -1, dtype=np.int64)
    first_use = np.full(N, n, dtype=np.int64)
    last_target = np.full(N, -1, dtype=np.int64)
    if n:
        np.maximum.at(last_use, bq, idx)
        np.minimum.at(first_use, bq, idx)
        np.maximum.at(last_target, bq, idx)
        cxs = bc >= 0
        np.maximum.at(last_use, bc[cxs], idx[cxs])
        np.minimum.at(first_use, bc[cxs], idx[cxs])
    # end of the maximal run of gates with the same target that starts at each gate
    run_end = np.empty(n + 1, dtype=np.int64)
    run_end[n] = n
    if n:
        change = np.flatnonzero(bq[1:] != bq[:-1]) + 1
        bounds = np.concatenate([change, [n]])
        run_end[:n] = bounds[np.searchsorted(bounds, idx, side='right')]
    pos_arr = np.full(max(N, 1), -1, dtype=np.int32)
    gate_at = _BasisView(bk, bq, bc, bp)

    def emit(res):
        nonlocal phase
        new_ops, touched_q, ph = res
        phase += ph
        ops.extend(new_ops)
        zero.difference_update(touched_q)

    def retired(res, j):
        new_ops = res[0]
        return len(new_ops) == 1 and new_ops[0].kind == 'mux' and last_use[new_ops[0].target] < j

    def lifetime_fits(t, i):
        e = int(last_use[t]) + 1
        qs = np.unique(np.concatenate([bq[i:e], bc[i:e]]))
        return len(qs) - (1 if len(qs) and qs[0] < 0 else 0) <= q_max

    doom: Optional[set] = None
    i = 0
    while i < n:
        t = int(bq[i])
        if (bk[i] != 4 and t in zero and (last_target[t] == i or not lifetime_fits(t, i))):
            B = gate_at(i).base_matrix()
            emit(([FusedOp('mux', t, (), np.array([B], dtype=np.complex128), True, 1)], [t], 0.0))
            i += 1
            continue
        if first_use[t] == i:
            j = int(run_end[i])
            if last_use[t] == j - 1:
                res = _run_mux([gate_at(g) for g in range(i, j)], t in zero)
                if res is not None:
                    emit(res)
                    i = j
                    continue
        blk = _Block(zero)
        best = None
        j = i
        doomed = doom is not None and sum(1 for q in doom if q in zero) >= 2
        live_i = int((last_use >= i).sum())                    # distinct qubits used by gates[i:]
        stopped = False
        while j < n:
            qs = (int(bq[j]),) if bc[j] < 0 else (int(bc[j]), int(bq[j]))
            new = [q for q in qs if q not in blk.pos]
            if new and blk.qubits:
                res = _classify(blk)
                if res is not None:
                    best = (j, res)
                    if retired(res, j):
                        stopped = True
                        break
                if len(blk.qubits) + len(new) > q_max:
                    stopped = True
                    break
            for q in new:
                blk.add_qubit(q)
                pos_arr[q] = blk.pos[q]
            U = blk.U
            if not U.flags.c_contiguous:
                U = blk.U = np.ascontiguousarray(U)
            # a doomed block stops right after the gate that completes its qubit set (see `doom` in fuse)
            limit = j + 1 if (doomed and len(blk.qubits) == live_i) else n
            j2 = int(apply_run(U.ctypes.data, U.shape[0], U.shape[1], bk.ctypes.data, bq.ctypes.data, bc.ctypes.data,
                               bp.ctypes.data, j, limit, pos_arr.ctypes.data))
            if j2 < 0:
                raise ValueError('basis program holds an unknown gate kind')
            blk.n_gates += j2 - j
            j = j2
            if doomed and len(blk.qubits) == live_i:
                stopped = True
                break
        for q in blk.qubits:
            pos_arr[q] = -1
        if not stopped:
            res = _classify(blk)
            if res is not None:
                best = (j, res)
            else:
                left = _unrestored_zero_qubits(blk)
                doom = left if len(left) >= 2 else None
        if best is None:
            best = (i + 1, _single_gate_op(gate_at(i), zero))
        emit(best[1])
        i = best[0]
    return _fold_init(prog.n_qubits, ops, phase, n0)


class _BasisView:
    """gate i of a basis-gate array program as an ir.Gate (built on demand: only single gates and short runs need it)."""

    def __init__(self, bk, bq, bc, bp):
        self.bk, self.bq, self.bc, self.bp = bk, bq, bc, bp

    def __call__(self, i):
        from .ir import BASIS_KINDS
        k, q = int(self.bk[i]), int(self.bq[i])
        if k == 0:
            return Gate('rz', (q,), (float(self.bp[i]),))
        if k == 4:
            return Gate('cx', (int(self.bc[i]), q), (), (1,))
        return Gate(BASIS_KINDS[k], (q,))


def _fold_init(n_qubits, ops, phase, n_gates_in) -> FusedCircuit:
    """Uncontrolled sweeps on qubits that nothing has touched yet belong to the product-state initialiser: the first on
    a |0> qubit sets its 2-vector, further ones (the rest of a transpiled H: the fusion loop peels rz.sx.rz gate by
    gate) multiply it."""
    init: Dict[int, np.ndarray] = {}
    kept: List[FusedOp] = []
    entangled: set = set()                         # qubits some kept sweep involves
    for op in ops:
        if op.kind == 'mux' and not op.ctrls and op.target not in entangled and (op.zero_in or op.target in init):
            v = init.get(op.target)
            init[op.target] = op.table[0][:, 0].copy() if v is None else op.table[0] @ v
        else:
            kept.append(op)
            entangled.update(op.ctrls)
            if op.kind == 'mux':
                entangled.add(op.target)
    return FusedCircuit(n_qubits, init, kept, phase, n_gates_in)


# ------------------------------------------------------------------------------------------
@dataclass
class Plan:
    """Engine program for one circuit."""
    n_logical: int
    n_phys: int                         # qubits the state buffer must hold
    layout: List[int]                   # logical qubit -> physical (>= n_phys: never stored, always 0)
    ops: np.ndarray                     # structured array matching qcm_op
    tables: np.ndarray                  # float64
    n_passes: int = 0
    n_sweeps_unblocked: int = 0
    bytes_algorithmic: int = 0          # per amplitude byte: filled by the backend
    final_active: int = 0
    global_phase: float = 0.0
    #: sharded layouts only: the n_global highest physical positions hold product-state qubits that
    #: are never targeted again (pure controls); position -> their 2-vector.  Ops cover positions
    #: below n_phys - n_global only.
    n_global: int = 0
    global_init: Dict[int, np.ndarray] = field(default_factory=dict)


OP_DTYPE = np.dtype([('kind', '<i4'), ('target', '<i4'), ('n_ctrl', '<i4'), ('n_active_in', '<i4'),
                     ('n_active_out', '<i4'), ('flags', '<i4'), ('ctrl', '<i4', (QCM_MAX_CTRL,)),
                     ('table_off', '<i8')], align=True)


class _Emitter:
    def __init__(self):
        self.ops = []
        self.tabs = []
        self.off = 0

    def table(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        if self.off % 2:                           # keep 16-byte alignment of fp64 pairs
            self.tabs.append(np.zeros(1))
            self.off += 1
        o = self.off
        self.tabs.append(arr)
        self.off += arr.size
        return o

    def op(self, kind, target=0, ctrl=(), n_in=0, n_out=0, table_off=0, n_ctrl=None):
        rec = np.zeros((), dtype=OP_DTYPE)
        rec['kind'], rec['target'] = kind, target
        rec['n_ctrl'] = len(ctrl) if n_ctrl is None else n_ctrl
        rec['n_active_in'], rec['n_active_out'] = n_in, n_out
        rec['table_off'] = table_off
        c = np.zeros(QCM_MAX_CTRL, dtype=np.int32)
        c[:len(ctrl)] = ctrl
        rec['ctrl'] = c
        self.ops.append(rec)

    def finish(self):
        ops = np.array(self.ops, dtype=OP_DTYPE) if self.ops else np.zeros(0, dtype=OP_DTYPE)
        tabs = np.concatenate(self.tabs) if self.tabs else np.zeros(0)
        return ops, tabs


def _mux_table_f64(table):
    t = np.empty((table.shape[0], 8))
    flat = table.reshape(-1, 4)
    t[:, 0::2] = flat.real
    t[:, 1::2] = flat.imag
    return t


def _diag_table_f64(table):
    t = np.empty((table.shape[0], 2))
    t[:, 0], t[:, 1] = table.real, table.imag
    return t


QCM_FLAG_SAMPLE_CHECKPOINT = 1
QCM_FLAG_ROTATED_OUTPUT_OK = 2          # the engine may store the last pass's result in its own address order


def _flag_last_pass(ops):
    """Set QCM_FLAG_SAMPLE_CHECKPOINT | QCM_FLAG_ROTATED_OUTPUT_OK on the header of the program's last
    pass: shots will be drawn from it, and nothing but post-selection / sampling / read-back follows, so
    a wide expansion may write its result as one sequential stream (include/qcmrf_b200.h)."""
    i, last = 0, -1
    while i < len(ops):
        last = i
        i += 1 + (int(ops[i]['n_ctrl']) if ops[i]['kind'] == QCM_OP_BLOCK else 0)
    if last >= 0 and ops[last]['kind'] in (QCM_OP_BLOCK, QCM_OP_MUX1Q):
        ops[last]['flags'] = QCM_FLAG_SAMPLE_CHECKPOINT | QCM_FLAG_ROTATED_OUTPUT_OK


def split_releasable(fc: FusedCircuit, keep_below: int = 0):
    """Measure-and-release: a qubit that is materialised from |0> by one
    sweep and never used again -- a QCMRF clique ancilla, measured right after its block
    (QCMRF.py:231-239) -- need not be stored at all.  Returns (core circuit without those sweeps,
    list of the released sweeps).  Qubits below ``keep_below`` (the variable register) are kept."""
    fc = drop_unmaterialised_controls(fc)
    used_later = {}
    last_targeted = {}
    for k, op in enumerate(fc.ops):
        for q in ((op.target,) if op.kind == 'mux' else ()) + tuple(op.ctrls):
            used_later[q] = k
        if op.kind == 'mux':
            last_targeted[op.target] = k
    core, virtual = [], []
    for k, op in enumerate(fc.ops):
        # the released outcome is drawn from (and its projection applied to) the FINAL value of the
        # sweep's index qubits: only exact if no later sweep changes them (non-diagonal = 'mux' target)
        if (op.kind == 'mux' and op.zero_in and op.target >= keep_below and used_later[op.target] == k
                and op.target not in fc.init and all(last_targeted.get(c, -1) < k for c in op.ctrls)):
            virtual.append(op)
        else:
            core.append(op)
    return FusedCircuit(fc.n_qubits, dict(fc.init), core, fc.global_phase, fc.n_gates_in), virtual


_MERGE_STEPS = {}


def _merge_steps(ctrls):
    """Structure of merge_diagonals for a list of index-qubit tuples: per member (flush?, union, gather
    index of the running table, gather index of the member's table) -- depends on the qubits only, so a
    sweep over one graph computes it once."""
    steps = _MERGE_STEPS.get(ctrls)
    if steps is None:
        steps, cur = [], []
        for ctrl in ctrls:
            ctrl = list(ctrl)
            union = list(cur) + [c for c in ctrl if c not in cur]
            flush = len(union) > QCM_MAX_CTRL
            if flush:
                cur, union = [], ctrl
            idx = np.arange(1 << len(union))
            a = np.zeros_like(idx)
            for j, c in enumerate(cur):
                a |= ((idx >> union.index(c)) & 1) << j
            b = np.zeros_like(idx)
            for j, c in enumerate(ctrl):
                b |= ((idx >> union.index(c)) & 1) << j
            steps.append((flush, tuple(union), a, b))
            cur = union
        if len(_MERGE_STEPS) > 256:
            _MERGE_STEPS.clear()
        _MERGE_STEPS[ctrls] = steps
    return steps


def merge_diagonals(members):
    """Product of diagonal factors [(ctrl positions, complex table 2^m)] over the union of their index
    qubits, at most QCM_MAX_CTRL bits per merged table.  Returns [(ctrl positions, complex table)]."""
    members = list(members)
    steps = _merge_steps(tuple(tuple(int(c) for c in ctrl) for ctrl, _d in members))
    out = []
    cur_ctrl, cur_tab = [], np.ones(1, dtype=np.complex128)
    for (flush, union, a, b), (_ctrl, d) in zip(steps, members):
        if flush:
            out.append((cur_ctrl, cur_tab))
            cur_tab = np.ones(1, dtype=np.complex128)
        cur_tab = cur_tab[a] * np.asarray(d)[b]
        cur_ctrl = list(union)
    out.append((cur_ctrl, cur_tab))
    return out


_SORT_IDX = {}


def sort_diag_ctrl(ctrl, tab):
    """The same diagonal factor with its index qubits in ascending order (the engine extracts runs of consecutive
    index qubits with one shift + mask each).  Works on a table with leading batch axes."""
    ctrl = tuple(int(c) for c in ctrl)
    order = sorted(range(len(ctrl)), key=lambda j: ctrl[j])
    if order == list(range(len(ctrl))):
        return list(ctrl), tab
    idx = _SORT_IDX.get(ctrl)
    if idx is None:
        k = np.arange(1 << len(ctrl))
        idx = np.zeros_like(k)
        for j, oj in enumerate(order):                        # new index bit j <-> old index bit order[j]
            idx |= ((k >> j) & 1) << oj
        if len(_SORT_IDX) > 256:
            _SORT_IDX.clear()
        _SORT_IDX[ctrl] = idx
    return [ctrl[j] for j in order], np.asarray(tab)[..., idx]


def control_only_qubits(fc: FusedCircuit) -> List[int]:
    """Product-state qubits that no later sweep targets: they only ever select table entries, so a
    state can be split on them across GPUs with no communication at all."""
    targeted = {op.target for op in fc.ops if op.kind == 'mux'}
    return [q for q in sorted(fc.init) if q not in targeted]


def plan(fc: FusedCircuit, lazy: bool = True, block_max: int = 4, elide: Optional[bool] = None,
         keep_order: bool = False, n_global: int = 0, expand_max: int = QCM_MAX_EXPAND) -> Plan:
    """Lay the fused circuit out for the engine.

    lazy=False  : identity layout, every qubit materialised up front, one pass per
                  fused op (the plain in-place "gate pass" execution, Aer-like width).
    lazy=True   : first-use layout, lazy materialisation, BLOCK passes of up to
                  ``block_max`` targets -- up to ``expand_max`` when every target of the pass is
                  a new qubit materialised by exactly one sweep (an "expansion" pass: wider is
                  better, the intermediate states all but vanish from the byte count); in a run
                  of such sweeps the FIRST pass takes the remainder so that the last, largest
                  state is written by a full-width pass.  Never-materialised qubits are not
                  stored unless elide=False.
    n_global=g  : (lazy only) the g highest-numbered control-only qubits are laid out on the g
                  highest physical positions, to be held by the rank index of a 2^g-way
                  sharded state; raises ValueError if the circuit has fewer than g of them.
    """
    fc = drop_unmaterialised_controls(fc)
    N = fc.n_qubits
    em = _Emitter()
    if elide is None:
        elide = lazy
    if not lazy:
        layout = list(range(N))
        order = list(range(N))
    else:
        order = sorted(fc.init.keys())
        seen = set(order)
        for op in fc.ops:
            for q in ((op.target,) if op.kind == 'mux' else ()) + tuple(op.ctrls):
                if q not in seen:
                    seen.add(q)
                    order.append(q)
        rest = [q for q in range(N) if q not in seen]
        gq: List[int] = []
        if n_global:
            cand = control_only_qubits(fc)
            if len(cand) < n_global:
                raise ValueError('circuit has %d control-only qubits, %d needed' % (len(cand), n_global))
            gq = cand[-n_global:]
            order = [q for q in order if q not in gq] + gq
        layout = [0] * N
        for p, q in enumerate(order):
            layout[q] = p
        for p, q in enumerate(rest):
            layout[q] = len(order) + p
        if not elide:
            order = order + rest
    n_phys = len(order)
    if n_global and (not lazy or not elide):
        raise ValueError('n_global needs the lazy, eliding layout')
    global_init = {}
    if n_global:
        for q in gq:
            global_init[layout[q]] = fc.init[q]

    # ---- INIT_PRODUCT ------------------------------------------------------------------
    if lazy:
        n_init = len(fc.init) - n_global
        if not elide and not fc.ops:
            n_init = n_phys
    else:
        n_init = n_phys
    qv = np.zeros((max(n_init, 1), 4))
    qv[:, 0] = 1.0
    for q, v in fc.init.items():
        p = layout[q]
        if p < n_init:
            qv[p] = [v[0].real, v[0].imag, v[1].real, v[1].imag]
    em.op(QCM_OP_INIT_PRODUCT, n_in=0, n_out=n_init, table_off=em.table(qv))
    active = n_init
    n_passes = 1

    # ---- sweeps ----------------------------------------------------------------------------
    pend: List[Tuple[int, Tuple[int, ...], int]] = []      # (phys target, phys ctrls, table_off)
    pend_targets: List[int] = []
    pend_in = active

    def flush():
        nonlocal pend, pend_targets, pend_in, active, n_passes
        if not pend:
            return
        tq = sorted(pend_targets)
        n_out = max(active, max(tq) + 1)
        if len(pend) == 1:
            t, c, off = pend[0]
            em.op(QCM_OP_MUX1Q, target=t, ctrl=c, n_in=pend_in, n_out=n_out, table_off=off)
        else:
            em.op(QCM_OP_BLOCK, target=len(tq), ctrl=tq, n_in=pend_in, n_out=n_out, n_ctrl=len(pend))
            for t, c, off in pend:
                em.op(QCM_OP_MUX1Q, target=t, ctrl=c, n_in=pend_in, n_out=n_out, table_off=off)
        active = n_out
        n_passes += 1
        pend, pend_targets = [], []
        pend_in = active

    # expansion sweeps: a mux whose target has not been stored or used before it (lazy layouts only)
    n_ops = len(fc.ops)
    is_exp = [False] * n_ops
    if lazy:
        seen_q = set(fc.init)
        for k, op in enumerate(fc.ops):
            if op.kind == 'mux' and op.zero_in and op.target not in seen_q:
                is_exp[k] = True
            seen_q.update(op.ctrls)
            if op.kind == 'mux':
                seen_q.add(op.target)
    run_left = [0] * (n_ops + 1)                   # expansion sweeps from k to the end of their run
    for k in range(n_ops - 1, -1, -1):
        run_left[k] = run_left[k + 1] + 1 if is_exp[k] else 0
    emax = max(block_max, min(expand_max, QCM_MAX_EXPAND)) if lazy else 1
    pend_exp = True                                # every pending member is an expansion sweep
    exp_limit = emax

    for k, op in enumerate(fc.ops):
        if op.kind == 'diag':
            flush()
            ctrl = tuple(layout[q] for q in op.ctrls)
            em.op(QCM_OP_DIAG, ctrl=ctrl, n_in=active, n_out=active, table_off=em.table(_diag_table_f64(op.table)))
            n_passes += 1
            continue
        t = layout[op.target]
        ctrl = tuple(layout[q] for q in op.ctrls)
        if len(ctrl) > QCM_MAX_CTRL:
            raise ValueError('fused op has %d index qubits (max %d)' % (len(ctrl), QCM_MAX_CTRL))
        if t >= active and t not in pend_targets:
            # first touch of a never-materialised qubit: first-use layout makes it the
            # next physical qubit after everything materialised or pending
            expected = active + sum(1 for x in pend_targets if x >= active)
            if t != expected:
                raise AssertionError('layout/materialisation order mismatch: target %d, expected %d' % (t, expected))
        new_targets = set(pend_targets) | {t}
        wide = pend_exp and is_exp[k] and t not in pend_targets
        if not lazy:
            limit = 1
        elif wide:
            limit = exp_limit
        else:
            limit = block_max
        conflict = (len(new_targets) > limit or len(pend) >= QCM_MAX_MEMBERS or
                    any(c in new_targets for c in ctrl) or
                    any(t in pc for _, pc, _ in pend) or
                    (pend and not wide and len(pend_targets) > block_max))
        if conflict:
            flush()
        if not pend:
            pend_in = active
            pend_exp = True
            # start of a pass inside a run of expansion sweeps: the first pass of the run takes the remainder
            exp_limit = emax
            if is_exp[k] and (k == 0 or not is_exp[k - 1]) and run_left[k] > emax and run_left[k] % emax:
                exp_limit = run_left[k] % emax
        pend_exp = pend_exp and is_exp[k] and t not in pend_targets
        off = em.table(_mux_table_f64(op.table))
        pend.append((t, ctrl, off))
        if t not in pend_targets:
            pend_targets.append(t)
    flush()
    if not lazy or not elide:
        if active < n_phys:
            em.op(QCM_OP_EXTEND, n_in=active, n_out=n_phys)
            active = n_phys
    ops, tabs = em.finish()
    # shots follow the program: let the engine build the sampler's tree before a final expansion pass
    _flag_last_pass(ops)
    return Plan(N, n_phys, layout, ops, tabs, n_passes, len(fc.ops) + 1, 0, active, fc.global_phase,
                n_global, global_init)
