def circuit_to_gate(circuit, **_):
    """Imported (unused) by QCMRF.py:8."""
    return circuit.to_gate()


def circuit_to_instruction(circuit, **_):
    return circuit.to_instruction()
