from qcmrf_b200.circuit import QuantumCircuit, Instruction, Gate, CircuitInstruction   # noqa: F401
