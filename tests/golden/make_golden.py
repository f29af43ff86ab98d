"""Regenerate tests/golden/*.json from the reference checkout (run in the build
container only; the GPU box has no /root/reference).

    python tests/golden/make_golden.py [/root/reference]

Writes
  models.json          {"0.1": {"GRAPHS":..., "THETAS":...}, "0.25":..., "0.5":...}
                       -- verbatim content of res_*/models*.json (data, not source)
  aer_counts.json      {"0.1": [70 dicts], ...} -- res_*/result_simulation.json, the
                       Aer qasm_simulator histograms (10000 shots, unseeded)
  quasi_sample.json    first 3 entries of res_0.1/result_torino.json (schema sample for
                       eval.py's 'quasi_dists' branch, eval.py:55-57)
  ref_programs.json    (only if the reference module imports under the compat shim)
                       gate programs emitted by the reference's own QCMRF.py for rep 0 of
                       each graph, flattened by qcmrf_b200.ir -- pins the product's
                       circuit constructor and the oracle's program against the
                       reference's _build (QCMRF.py:199-243).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'


def main():
    models, counts = {}, {}
    for scale, mfile in (('0.1', 'models_0.1.json'), ('0.25', 'models_0.25.json'),
                         ('0.5', 'models.json')):
        d = os.path.join(REF, 'res_' + scale)
        models[scale] = json.load(open(os.path.join(d, mfile)))
        counts[scale] = json.load(open(os.path.join(d, 'result_simulation.json')))
    json.dump(models, open(os.path.join(HERE, 'models.json'), 'w'), separators=(',', ':'))
    json.dump(counts, open(os.path.join(HERE, 'aer_counts.json'), 'w'), separators=(',', ':'))
    t = json.load(open(os.path.join(REF, 'res_0.1', 'result_torino.json')))
    json.dump({'quasi_dists': t['quasi_dists'][:3], 'metadata': t['metadata'][:3]},
              open(os.path.join(HERE, 'quasi_sample.json'), 'w'), separators=(',', ':'))

    # Programs from the reference's own constructor, run under the compat shim.
    repo = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, repo)
    try:
        from qcmrf_b200 import compat
        compat.install()
        sys.path.insert(0, REF)
        import QCMRF as ref_mod          # the reference module itself
        from qcmrf_b200 import ir
    except Exception as e:               # shim not built yet
        print('ref_programs.json skipped:', e)
        return
    progs = {}
    for scale in models:
        for j, C in enumerate(models[scale]['GRAPHS']):
            theta = models[scale]['THETAS'][str(j)][0]
            for wm in (True, False):
                circ = ref_mod.QCMRF(C, theta, with_measurements=wm)
                prog = ir.lower(circ)
                progs['%s/%d/%d' % (scale, j, int(wm))] = ir.to_jsonable(prog)
    json.dump(progs, open(os.path.join(HERE, 'ref_programs.json'), 'w'), separators=(',', ':'))
    print('wrote', len(progs), 'reference programs')


if __name__ == '__main__':
    main()
