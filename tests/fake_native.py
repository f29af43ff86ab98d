"""CPU stand-in for the CUDA engine, for HOST-LOGIC tests only.

TEST INFRASTRUCTURE: executes engine programs with tests/engine_emulator.py (numpy)
so that the lowering / fusion / planning / key-formatting code -- and the
reference's own scripts on top of it -- can be exercised in the GPU-less build
container.  It is monkeypatched over qcmrf_b200._native by the tests that need it;
the product has no such path (without the CUDA library it raises).
"""
import numpy as np

import engine_emulator as em


class FakePlanHolder:
    pass


def _keys_from_probs(p, shots, rng, clbit_map):
    idx = rng.choice(len(p), size=shots, p=p / p.sum())
    if clbit_map is None or len(clbit_map) == 0:
        return idx.astype(np.uint64)
    keys = np.zeros(shots, dtype=np.uint64)
    for c, q in enumerate(clbit_map):
        if q >= 0:
            keys |= ((idx >> int(q)) & 1).astype(np.uint64) << np.uint64(c)
    return keys


def small_max_qubits(precision='double'):
    return 13


def run_batch_small(plans, clbit_maps, ps, shots, seed, precision='double', device=0, want_probs=True,
                    stream_ids=None):
    n = len(plans)
    keys = np.zeros((n, shots), dtype=np.uint64) if shots else None
    probs, kept = [], np.zeros(n)
    for i, pl in enumerate(plans):
        phys, act = em.run_plan(pl)
        w = np.abs(phys) ** 2
        mask, value, bits = ps[i]
        idx = np.arange(len(w))
        sel = (idx & mask) == value
        out = np.zeros(1 << bits)
        np.add.at(out, idx[sel] & ((1 << bits) - 1), w[sel])
        probs.append(out)
        kept[i] = w[sel].sum()
        if shots:
            sid = i if stream_ids is None else int(stream_ids[i])
            rng = np.random.default_rng([seed & 0xffffffff, sid])
            keys[i] = _keys_from_probs(w, shots, rng, clbit_maps[i])
    return keys, probs, kept, 0.0


class Handle:
    def __init__(self, n_local, precision='single', device=0, ext_state_ptr=None, ext_stream=None):
        self.n_local = n_local
        self.cdtype = np.complex64 if precision in ('single', 'c64', 32) else np.complex128
        self.state = np.zeros(1 << n_local, dtype=np.complex128)
        self.active = 0

    def run_program(self, ops, tables):
        pl = FakePlanHolder()
        pl.ops, pl.tables, pl.n_phys = ops, tables, self.n_local
        self.state, self.active = em.run_plan(pl)

    def postselect(self, mask, value, n_out_bits, want_probs=True):
        w = np.abs(self.state) ** 2
        idx = np.arange(len(w))
        sel = (idx & mask) == value
        out = np.zeros(1 << n_out_bits)
        np.add.at(out, idx[sel] & ((1 << n_out_bits) - 1), w[sel])
        return out, float(w[sel].sum())

    def sample(self, shots, seed, stream_id=0, clbit_qubit=None):
        rng = np.random.default_rng([seed & 0xffffffff, stream_id])
        return _keys_from_probs(np.abs(self.state) ** 2, shots, rng, clbit_qubit)

    def get_amplitudes(self, first=0, count=None):
        count = len(self.state) - first if count is None else count
        return self.state[first:first + count].astype(self.cdtype)

    def timing(self):
        return dict(program_ms=0.0, sample_ms=0.0, postselect_ms=0.0, kernel_launches=0, bytes_read=0,
                    bytes_written=0)

    def close(self):
        pass


def install(monkeypatch):
    from qcmrf_b200 import _native
    monkeypatch.setattr(_native, 'small_max_qubits', small_max_qubits)
    monkeypatch.setattr(_native, 'run_batch_small', run_batch_small)
    monkeypatch.setattr(_native, 'Handle', Handle)
