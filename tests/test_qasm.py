"""OpenQASM 2 in/out (host logic): round trips and a hand-written foreign circuit vs the oracle."""
import numpy as np

from oracle import program, statevector as sv
from qcmrf_b200 import QCMRF, ir, qasm, transpile


def _state(prog):
    psi, _ = sv.run_program(ir.to_oracle_ops(prog), prog.n_qubits)
    return psi


def test_roundtrip_untranspiled_and_transpiled(models):
    C = models['0.5']['GRAPHS'][2]
    th = models['0.5']['THETAS']['2'][3]
    for circ in (QCMRF(C, th), transpile([QCMRF(C, th)], basis_gates=['cx', 'id', 'rz', 'sx', 'x'])[0]):
        p0 = ir.lower(circ)
        text = qasm.dumps(circ)
        assert text.startswith('OPENQASM 2.0;')
        p1 = qasm.loads(text)
        assert (p1.n_qubits, p1.n_clbits, p1.measures) == (p0.n_qubits, p0.n_clbits, p0.measures)
        a, b = _state(p0), _state(p1)
        assert np.abs(a - b * np.exp(1j * (p0.global_phase - p1.global_phase))).max() < 1e-13
        assert qasm.dumps(p1) == qasm.dumps(qasm.loads(qasm.dumps(p1)))          # fixed point


def test_foreign_circuit_with_registers_and_expressions():
    text = '''OPENQASM 2.0;
    include "qelib1.inc";
    qreg a[2]; qreg b[1];
    creg m[3];
    h a;                       // register broadcast
    u3(pi/2, 0, pi) b[0];
    cx a[0], b[0];
    cu1(pi/4) a[1], b[0];
    ccx a[0], a[1], b[0];
    rz(-pi/8) a[1]; sx a[0]; swap a[0], a[1];
    barrier a, b;
    measure a -> m;            // clbits 0,1 <- qubits 0,1
    measure b[0] -> m[2];
    '''
    p = qasm.loads(text)
    assert p.n_qubits == 3 and p.n_clbits == 3 and p.measures == {0: 0, 1: 1, 2: 2}
    names = [g.name for g in p.gates]
    assert names == ['h', 'h', 'u', 'cx', 'cp', 'mcx', 'rz', 'sx', 'cx', 'cx', 'cx']
    # against an independent numpy construction
    from qcmrf_b200.ir import one_qubit_matrix
    psi = np.zeros(8, dtype=complex); psi[0] = 1
    for g in p.gates:
        U = g.matrix()
        full = np.zeros((8, 8), dtype=complex)
        for i in range(8):
            for j in range(8):
                rest_i = [(i >> q) & 1 for q in range(3) if q not in g.qubits]
                rest_j = [(j >> q) & 1 for q in range(3) if q not in g.qubits]
                if rest_i != rest_j:
                    continue
                li = sum(((i >> q) & 1) << k for k, q in enumerate(g.qubits))
                lj = sum(((j >> q) & 1) << k for k, q in enumerate(g.qubits))
                full[i, j] = U[li, lj]
        psi = full @ psi
    from qcmrf_b200 import fusion
    import engine_emulator as em
    fc = fusion.fuse(p, 'clique')
    pl = fusion.plan(fc, lazy=True, block_max=2)
    phys, _ = em.run_plan(pl)
    got = em.logical_state(pl, phys)
    assert np.abs(got - psi).max() < 1e-12
