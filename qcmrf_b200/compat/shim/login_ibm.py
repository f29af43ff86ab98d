"""Stand-in for the private credentials module the reference imports
(run_experiment.py:13, eval.py:8); only reachable from dead code."""


def service(*a, **k):
    raise RuntimeError('login_ibm.service: no IBM Quantum credentials in this environment')
