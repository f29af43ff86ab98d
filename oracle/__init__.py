"""CPU oracle for the QCMRF statevector hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import or execute it, and only as the checker or
as the timed CPU baseline -- never as a fallback for the CUDA path.

What it restates (reference = np84/qcmrf, mounted at /root/reference):

* ``oracle.program``      -- the gate program ``QCMRF._build`` emits
                             (QCMRF.py:199-243), theta->gamma (QCMRF.py:144-157).
* ``oracle.statevector``  -- gate-by-gate complex128 statevector execution of
                             that program plus multinomial shot sampling.  The
                             reference delegates this to qiskit-aer's
                             ``qasm_simulator`` (run_experiment.py:54-57), an
                             un-vendored, un-pinned third-party dependency
                             (qiskit<1.0 / qiskit-aer ~0.13 by inference from
                             the imports); its published
                             algorithm (dense little-endian statevector, one
                             sweep per gate, deferred measurement sampling
                             from |psi|^2) is what is restated here.
* ``oracle.mrf``          -- brute-force MRF enumeration, standing in for the
                             proprietary ``kiopto_native`` exact inference the
                             reference's eval.py uses (eval.py:84-93), and the
                             post-selection / fidelity arithmetic
                             (QCMRF.py:247-284, eval.py:115-123).
* ``oracle/csrc``         -- the same executor in plain C + OpenMP; it is the
                             timed CPU baseline ("port").

Parity pin: the reference ships no exact vectors for this path.  The oracle is
pinned against everything it does ship (tests/test_oracle_golden.py):
the three ``models*.json`` files (regenerated bit-exactly from seed 1984), the
210 Aer histograms in ``res_*/result_simulation.json`` (unseeded => statistical
pin: chi^2 / TV / success rate), and the bit conventions those keys imply.
Aer's exact statevector and the transpiler's output are NOT recorded anywhere
in the reference: for those, parity is unpinned.
"""
