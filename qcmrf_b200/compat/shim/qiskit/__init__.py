"""qiskit stand-in (only the surface np84/qcmrf touches) backed by qcmrf_b200."""
from qcmrf_b200.circuit import QuantumCircuit, Instruction, Gate   # noqa: F401
from qcmrf_b200.transpile import transpile                          # noqa: F401

__version__ = '0.0-qcmrf_b200-shim'
__qcmrf_b200_shim__ = True


class _AerProvider:
    """``from qiskit import Aer`` -> ``Aer.get_backend('qasm_simulator')``
    (run_experiment.py:14,54)."""

    _NAMES = ('qasm_simulator', 'aer_simulator', 'statevector_simulator', 'aer_simulator_statevector')

    def get_backend(self, name='qasm_simulator', **options):
        if name not in self._NAMES:
            raise LookupError("backend %r not available (have %s)" % (name, ', '.join(self._NAMES)))
        from qcmrf_b200.backend import B200Simulator
        return B200Simulator(name=name, **options)

    def backends(self):
        return list(self._NAMES)


Aer = _AerProvider()


def execute(circuits, backend, shots=1024, **kw):
    return backend.run(circuits, shots=shots, **kw)
