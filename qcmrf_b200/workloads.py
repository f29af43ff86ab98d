"""The synthetic MRFs BASELINE.json's configs name, seeded.

Every vertex 0..n-1 appears in some clique because the constructor infers n from the
largest index (/root/reference/QCMRF.py:52-57); theta follows the reference's prior,
``-halfnorm(scale)`` (/root/reference/run_experiment.py:30), so theta <= 0 and
cos 2gamma = exp(beta*theta/2) is in (0, 1].
"""
import numpy as np

__all__ = ['random_tree', 'chain', 'theta_for', 'named']


def random_tree(n, extra_edges=0, seed=1984):
    """Random spanning tree on n vertices (edge (rand(0..v-1), v) for v = 1..n-1) plus
    ``extra_edges`` distinct extra random edges."""
    rng = np.random.RandomState(seed)
    cliques = [[int(rng.randint(0, v)), v] for v in range(1, n)]
    have = {tuple(c) for c in cliques}
    while extra_edges > 0:
        a, b = sorted(int(x) for x in rng.choice(n, 2, replace=False))
        if (a, b) not in have:
            have.add((a, b))
            cliques.append([a, b])
            extra_edges -= 1
    return cliques


def chain(n):
    return [[i, i + 1] for i in range(n - 1)]


def theta_for(cliques, scale=0.5, seed=1984):
    """-|N(0, scale)| per clique state, from a private RandomState (same law as
    scipy.stats.halfnorm.rvs(scale=...) in run_experiment.py:30)."""
    rng = np.random.RandomState(seed)
    dim = sum(2 ** len(c) for c in cliques)
    return list(-np.abs(rng.standard_normal(dim)) * scale)


#: name -> (cliques, total qubits N = n + k + 1)
def named(name):
    if name == 'q33':                      # config 3: 33 total qubits, n=16, tree + 1 edge
        c = random_tree(16, 1)
    elif name == 'q34':                    # config 4: 34 total qubits, n=17, tree
        c = random_tree(17, 0)
    elif name == 'q35':                    # the largest QCMRF one B200 holds: 34 stored qubits = 128 GiB complex64
        c = random_tree(17, 1)
    elif name == 'q37':                    # config 4: 37 total qubits, n=18, tree + 1 edge
        c = random_tree(18, 1)
    elif name == 'chain20':                # config 2: 20-variable chain
        c = chain(20)
    elif name.startswith('tree'):          # treeNN: NN variables, N = 2*NN total qubits
        c = random_tree(int(name[4:]), 0)
    else:
        raise ValueError('unknown workload %r' % name)
    n = max(max(x) for x in c) + 1
    return c, n + len(c) + 1
