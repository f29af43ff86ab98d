"""Names imported by run_experiment.py:12 / eval.py:7.  The IBM-hardware branch of the
reference is dead code behind exit() (run_experiment.py:63-82) and needs cloud
credentials; these placeholders only make the import succeed."""


def _unavailable(name):
    class _Stub:
        def __init__(self, *a, **k):
            raise RuntimeError('%s: IBM Quantum runtime is not available in this environment' % name)
    _Stub.__name__ = name
    return _Stub


QiskitRuntimeService = _unavailable('QiskitRuntimeService')
Session = _unavailable('Session')
Estimator = _unavailable('Estimator')
Sampler = _unavailable('Sampler')


class Options:
    def __init__(self):
        class _NS:
            pass
        self.execution = _NS()
        self.resilience_level = 0
        self.optimization_level = 0
