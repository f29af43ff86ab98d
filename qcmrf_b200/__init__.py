"""qcmrf_b200 -- B200-native statevector simulator for QCMRF circuits.

Public surface:
    QCMRF, extract_probs, fidelity, KL      (mirror of the reference's QCMRF.py)
    QuantumCircuit, AND, transpile          (circuit layer / basis translation)
    B200Simulator, get_backend              (Aer-style backend over the CUDA engine)
    ExactMRF                                (exact MRF inference by GPU enumeration: eval.py's ground truth)
"""
from .circuit import AND, QuantumCircuit
from .mrf import KL, QCMRF, extract_probs, fidelity
from .transpile import transpile
from .backend import B200Simulator, Counts, Job, Result
from .exact import ExactMRF
from . import qasm, workloads

__version__ = '0.1.0'


def get_backend(name='qasm_simulator', **options):
    return B200Simulator(name=name, **options)
