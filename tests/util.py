'''Shared helpers for the test-suite: golden loading and Problem construction.'''

import glob
import os
import sys
import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from nutils_b200 import bspline, points  # noqa: E402
from oracle import fem_oracle  # noqa: E402


def golden_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, '*.npz')))


def load_golden(name):
    with numpy.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def bases_1d(nelems, degree, btype):
    cont = {'spline': -1, 'std': 0}[btype]
    return [bspline.spline_basis_1d(n, degree, continuity=cont) for n in nelems]


def problem_from_golden(g):
    nelems = tuple(int(n) for n in g['nelems'])
    degree = int(g['degree'])
    ndims = len(nelems)
    b1 = bases_1d(nelems, degree, str(g['btype']))
    rules = points.tensor_gauss(ndims, int(g['qdegree']))
    ncomp = ndims if str(g['kind']) == 'elasticity' else 1
    return fem_oracle.Problem(nelems, [degree] * ndims, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], g['nodes'], ncomp=ncomp)


def relerr(a, b):
    a = numpy.asarray(a)
    b = numpy.asarray(b)
    n = numpy.linalg.norm(b)
    return numpy.linalg.norm(a - b) / (n if n else 1.)


def rowsum_relerr(va, vb, rowptr):
    'max over rows of |rowsum(a)-rowsum(b)| relative to the largest absolute row sum of b'
    if len(va) == 0:
        return 0.
    starts = numpy.asarray(rowptr[:-1])
    nonempty = numpy.diff(rowptr) > 0
    ra = numpy.add.reduceat(va, starts[nonempty])
    rb = numpy.add.reduceat(vb, starts[nonempty])
    scale = numpy.add.reduceat(abs(vb), starts[nonempty]).max()
    return abs(ra - rb).max() / scale
