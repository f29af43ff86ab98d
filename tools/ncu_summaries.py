"""Turn the scratch ncu outputs under gpurun_out/ into the summaries committed under profiles/.
    python tools/ncu_summaries.py r01
- gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum --csv)  -> profiles/<tag>_launches_summary.csv,
  profiles/<tag>_launches_first40.csv
- gpurun_out/prof_*.ncu-rep (ncu --set full)                            -> profiles/<tag>_ncu_prof_*.csv (metric,unit,value)
"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')


def launches(tag, name='launches.csv', out=None):
    path = os.path.join(OUT, name)
    if not os.path.exists(path):
        return
    out = out or tag
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        ms = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[iu], 1e-6)
        seq.append((r[ik], ms))
    tot = {}
    for k, ms in seq:
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + ms)
    total = sum(t for _n, t in tot.values())
    with open(os.path.join(PROF, out + '_launches_summary.csv'), 'w') as f:
        f.write('# ncu launch list: python bench.py --steps 2 --warmup 3 --no-cpu-baseline  (gpu__time_duration.sum, '
                '--clock-control none; cold-cache, serialised: compare shares)\n')
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_ms', 'share'])
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, '%.3f' % t, '%.4f' % (t / total)])
    with open(os.path.join(PROF, out + '_launches_first40.csv'), 'w') as f:
        w = csv.writer(f)
        w.writerow(['id', 'kernel', 'ms'])
        for i, (k, ms) in enumerate(seq[:40]):
            w.writerow([i, k, '%.4f' % ms])


KEEP = ('dram__', 'gpu__time', 'launch__', 'lts__t_bytes', 'lts__throughput', 'l1tex__throughput',
        'l1tex__data_bank_conflicts', 'sm__throughput', 'sm__warps_active', 'smsp__issue_active',
        'smsp__inst_executed.sum', 'issue_stalled', 'sm__inst_executed_pipe', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.avg ', 'lts__t_sectors_srcunit_tex_op')


def full(tag):
    """gpurun_out/prof_*.ncu-rep and gpurun_out/<tag>_prof_*.ncu-rep -> profiles/<tag>_ncu_prof_<name>[_<kernel>].csv, one
    file per profiled kernel launch of the report."""
    for fn in sorted(os.listdir(OUT)):
        if not fn.endswith('.ncu-rep') or not fn.startswith(tag + '_prof_' if tag != 'r01' else 'prof_'):
            continue
        base = fn[:-len('.ncu-rep')]
        if base.startswith(tag + '_'):
            base = base[len(tag) + 1:]
        raw = subprocess.run(['ncu', '-i', os.path.join(OUT, fn), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ik = hdr.index('Kernel Name')
        for n, vals in enumerate(rows[2:]):
            suffix = ''
            if len(rows) > 3:
                kn = vals[ik].split('<')[0].split('(')[0].replace('void ', '').replace('qcm::', '').strip()
                suffix = '_' + kn
            with open(os.path.join(PROF, '%s_ncu_%s%s.csv' % (tag, base, suffix)), 'w') as f:
                w = csv.writer(f)
                w.writerow(['metric', 'unit', 'value'])
                for h, u, v in zip(hdr, units, vals):
                    if h in ('ID', 'Process ID', 'Process Name', 'Host Name', 'Context', 'Stream', 'Device', 'CC'):
                        continue
                    if '.' in h and not any(s in h for s in KEEP):
                        continue
                    w.writerow([h, u, v])


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    launches(tag, 'launches.csv' if tag == 'r01' else tag + '_launches.csv')
    for fn in sorted(os.listdir(OUT)):                       # e.g. gpurun_out/r02_chain20_launches.csv
        if fn.startswith(tag + '_') and fn.endswith('_launches.csv'):
            launches(tag, fn, fn[:-len('_launches.csv')])
    full(tag)
