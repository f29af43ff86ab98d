from qcmrf_b200.circuit import AND   # noqa: F401
