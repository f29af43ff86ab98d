"""Import shims that let the reference's scripts run unchanged on this backend.

/root/reference/run_experiment.py and eval.py import ``qiskit`` (QuantumCircuit,
transpile, Aer, opflow, converters, circuit.library.AND), ``qiskit_ibm_runtime``,
``login_ibm``, ``kiopto_native`` and ``prettytable``; none of them is installed in
this image and none can be (no network).  ``install()`` puts ``compat/shim`` on
``sys.path`` for exactly the names that are missing, so a real Qiskit, if present,
always wins.  ``Aer.get_backend('qasm_simulator')`` then returns the B200 backend.

Usage:  python -c "import qcmrf_b200.compat as c; c.install(); import runpy; runpy.run_path('run_experiment.py')"
   or:  PYTHONPATH=<repo>/qcmrf_b200/compat/shim:<repo> python run_experiment.py
"""
import importlib.util
import os
import sys

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')
SHIMMED = ('qiskit', 'qiskit_ibm_runtime', 'login_ibm', 'kiopto_native', 'prettytable')


def missing():
    out = []
    for name in SHIMMED:
        try:
            spec = importlib.util.find_spec(name)
        except (ImportError, ValueError):
            spec = None
        if spec is None or (spec.origin or '').startswith(SHIM_DIR):
            out.append(name)
    return out


def install(force=False):
    """Make the shimmed names importable.  Returns the list of names served by the shim."""
    names = list(SHIMMED) if force else missing()
    if names and SHIM_DIR not in sys.path:
        if force:
            sys.path.insert(0, SHIM_DIR)
        else:
            sys.path.append(SHIM_DIR)      # real packages, if any, keep precedence
    return names
