"""OpenQASM 2.0 in and out for the engine's gate program (foreign circuits run too,
and transpiled fixtures can be stored as text).

``loads`` reads the qelib1 subset a basis-gate circuit uses -- ``id x y z h s sdg t tdg sx sxdg rz rx
ry p u1 u2 u3 u cx cz cp cu1 crz ccx c3x c4x swap measure barrier`` over any number of ``qreg`` /
``creg`` (registers are concatenated in declaration order, as Qiskit numbers them) -- into an
``ir.Program``.  ``dumps`` writes a Program back; multi-controlled gates with open controls or more
controls than qelib1 offers are written as X-conjugated ``mcx_<k>`` / ``mcp_<k>`` lines, an extension
``loads`` understands (marked in the header).  Parameters are written with 17 significant digits, so
a round trip is exact.
"""
import math
import re

from .ir import Gate, Program, lower

__all__ = ['loads', 'dumps']

_ONE = {'id', 'x', 'y', 'z', 'h', 's', 'sdg', 't', 'tdg', 'sx', 'sxdg'}
_ONE_P = {'rz': 'rz', 'rx': 'rx', 'ry': 'ry', 'p': 'p', 'u1': 'p'}
_CTRL = {'cx': ('cx', 0), 'cz': ('cz', 0), 'cy': ('cy', 0), 'ch': ('ch', 0), 'cp': ('cp', 1), 'cu1': ('cp', 1),
         'crz': ('crz', 1), 'crx': ('crx', 1), 'cry': ('cry', 1), 'csx': ('csx', 0),
         'ccx': ('mcx', 0), 'c3x': ('mcx', 0), 'c4x': ('mcx', 0)}
_SAFE = {'pi': math.pi, 'sin': math.sin, 'cos': math.cos, 'tan': math.tan, 'exp': math.exp, 'ln': math.log,
         'sqrt': math.sqrt}


def _num(expr):
    expr = expr.strip()
    if not re.fullmatch(r'[0-9eE\.\+\-\*/\(\)\s,a-z]*', expr):
        raise ValueError('bad parameter expression %r' % expr)
    return float(eval(expr.replace('^', '**'), {'__builtins__': {}}, _SAFE))          # arithmetic on numbers and pi only


def loads(text: str) -> Program:
    text = re.sub(r'//[^\n]*', '', text)
    stmts = [s.strip() for s in text.split(';') if s.strip()]
    qoff, coff, nq, nc = {}, {}, 0, 0
    prog = Program(0, 0)

    def qarg(tok):
        m = re.fullmatch(r'(\w+)\s*\[\s*(\d+)\s*\]', tok.strip())
        if m:
            return [qoff[m.group(1)][0] + int(m.group(2))]
        base, size = qoff[tok.strip()]
        return [base + i for i in range(size)]

    def carg(tok):
        m = re.fullmatch(r'(\w+)\s*\[\s*(\d+)\s*\]', tok.strip())
        if m:
            return [coff[m.group(1)][0] + int(m.group(2))]
        base, size = coff[tok.strip()]
        return [base + i for i in range(size)]

    for st in stmts:
        if st.startswith('OPENQASM') or st.startswith('include'):
            continue
        m = re.fullmatch(r'(qreg|creg)\s+(\w+)\s*\[\s*(\d+)\s*\]', st)
        if m:
            kind, name, size = m.group(1), m.group(2), int(m.group(3))
            if kind == 'qreg':
                qoff[name] = (nq, size)
                nq += size
            else:
                coff[name] = (nc, size)
                nc += size
            continue
        m = re.fullmatch(r'measure\s+(.+?)\s*->\s*(.+)', st)
        if m:
            for q, c in zip(qarg(m.group(1)), carg(m.group(2))):
                prog.measures[c] = q
            continue
        m = re.fullmatch(r'(\w+)\s*(?:\((.*)\))?\s+(.+)', st, flags=re.S)
        if not m:
            raise ValueError('cannot parse QASM statement %r' % st)
        name, params, args = m.group(1).lower(), m.group(2), m.group(3)
        if name == 'barrier':
            continue
        ps = [_num(p) for p in params.split(',')] if params else []
        operands = [qarg(a) for a in args.split(',')]
        width = max(len(o) for o in operands)
        for k in range(width):                                       # register broadcast
            qs = tuple(o[k] if len(o) > 1 else o[0] for o in operands)
            _emit(prog, name, ps, qs)
    prog.n_qubits, prog.n_clbits = nq, nc
    return prog


def _emit(prog, name, ps, qs):
    g = prog.gates
    if name in _ONE:
        g.append(Gate(name, qs))
    elif name in _ONE_P:
        g.append(Gate(_ONE_P[name], qs, (ps[0],)))
    elif name == 'u2':
        g.append(Gate('u', qs, (math.pi / 2, ps[0], ps[1])))
    elif name in ('u3', 'u'):
        g.append(Gate('u', qs, tuple(ps)))
    elif name in _CTRL:
        canon, npar = _CTRL[name]
        g.append(Gate(canon, qs, tuple(ps[:npar]), (1,) * (len(qs) - 1)))
    elif name == 'swap':
        a, b = qs
        for c, t in ((a, b), (b, a), (a, b)):
            g.append(Gate('cx', (c, t), (), (1,)))
    elif re.fullmatch(r'mc[xp]_\d+', name):                          # qcmrf_b200 extension written by dumps()
        g.append(Gate('mcx' if name[2] == 'x' else 'mcp', qs, tuple(ps), (1,) * (len(qs) - 1)))
    else:
        raise ValueError('qcmrf_b200.qasm: unsupported gate %r' % name)


def _f(x):
    return repr(float(x))


def dumps(circuit) -> str:
    prog = lower(circuit)
    out = ['OPENQASM 2.0;', 'include "qelib1.inc";',
           '// mcx_<k> / mcp_<k>(lambda): k-controlled X / phase, controls first -- qcmrf_b200 extension',
           'qreg q[%d];' % prog.n_qubits]
    if prog.n_clbits:
        out.append('creg c[%d];' % prog.n_clbits)
    if prog.global_phase:
        out.append('// global phase %s' % _f(prog.global_phase))
    for g in prog.gates:
        qs = ','.join('q[%d]' % q for q in g.qubits)
        nctrl = len(g.qubits) - 1
        if nctrl == 0:
            if g.name == 'u':
                out.append('u3(%s) %s;' % (','.join(_f(p) for p in g.params), qs))
            elif g.params:
                out.append('%s(%s) %s;' % (g.name, _f(g.params[0]), qs))
            else:
                out.append('%s %s;' % (g.name, qs))
            continue
        opened = [q for q, v in zip(g.controls, g.ctrl_values) if v == 0]
        for q in opened:
            out.append('x q[%d];' % q)
        if g.name in ('cx', 'mcx'):
            std = {1: 'cx', 2: 'ccx', 3: 'c3x', 4: 'c4x'}.get(nctrl)
            out.append('%s %s;' % (std or 'mcx_%d' % nctrl, qs))
        elif g.name in ('cp', 'mcp'):
            out.append(('cp(%s) %s;' if nctrl == 1 else 'mcp_' + str(nctrl) + '(%s) %s;') % (_f(g.params[0]), qs))
        elif nctrl == 1 and g.name in ('cz', 'cy', 'ch', 'csx'):
            out.append('%s %s;' % (g.name, qs))
        elif nctrl == 1 and g.name in ('crz', 'crx', 'cry'):
            out.append('%s(%s) %s;' % (g.name, _f(g.params[0]), qs))
        else:
            raise ValueError('qcmrf_b200.qasm: cannot write gate %r with %d controls' % (g.name, nctrl))
        for q in opened:
            out.append('x q[%d];' % q)
    for c, q in sorted(prog.measures.items()):
        out.append('measure q[%d] -> c[%d];' % (q, c))
    return '\n'.join(out) + '\n'
