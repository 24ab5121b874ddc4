'''Access to the UNMODIFIED reference installed under baseline/_ref (scripts/install_reference.py; git-ignored, travels to
the GPU box) and an oracle-backed stand-in for the device backend of nutils_b200.hook, so that the hook's plumbing can be
tested where there is no GPU.  TEST INFRASTRUCTURE.'''

import os
import sys
import numpy

from tests import util

REFDIR = os.path.join(util.ROOT, 'baseline', '_ref')


def have_reference():
    return os.path.isdir(os.path.join(REFDIR, 'nutils'))


def reference():
    'import the reference package (returns the nutils module)'
    for path in (os.path.join(REFDIR, 'examples'), REFDIR):
        if path not in sys.path:
            sys.path.insert(0, path)
    os.environ.setdefault('NUTILS_MATRIX', 'scipy')
    import nutils
    from nutils import export
    export.triplot = lambda *args, **kwargs: None   # matplotlib is absent; the examples only plot with it
    return nutils


class OracleBackend:
    'what hook.GpuBackend does, computed by the numpy oracle (checker for the host-side logic of the hook)'

    def __init__(self):
        self.calls = 0

    def _problem(self, spec):
        from oracle import fem_oracle
        b1, rules = spec['bases'], spec['rules']
        kw = {}
        if 'face' in spec:   # a side of the topology: the adjacent elements as an element set with points on their faces
            from nutils_b200 import points
            ids, xi, w = spec['elem_ids'], spec['xi'], spec['weights']
            kw = dict(elem_ids=ids, qoff=numpy.arange(len(ids) + 1) * len(w), qcoords=numpy.tile(xi, (len(ids), 1)), qweights=numpy.tile(w, len(ids)),
                      face_dim=numpy.full(len(ids), spec['face']['dim'], dtype=numpy.int8))
            rules = points.tensor_gauss(len(b1), 2)
        if 'elemset' in spec:   # trimmed / subset topology: kept elements, ragged points, pruned numbering
            from nutils_b200 import points
            kw = dict(spec['elemset'])
            if spec['renumber'] is not None:
                kw.update(renumber=spec['renumber'], nbasis_new=spec['nbasis_new'])
            rules = points.tensor_gauss(len(b1), 2)
        return fem_oracle.Problem(tuple(b.nelems for b in b1), [b.degree for b in b1], [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                                  [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], spec['nodes'], ncomp=spec['ncomp'], **kw)

    def pattern(self, spec):
        from oracle import fem_oracle
        na, nc = len(spec['bases']) + 1, spec['ncomp']
        mats, vecs = fem_oracle.assemble(self._problem(spec), [('generic', numpy.zeros((nc, na, nc, na)))], [])
        return mats[0][1], mats[0][2]

    def assemble(self, spec, Ds, Cs):
        from oracle import fem_oracle
        self.calls += 1
        mats, vecs = fem_oracle.assemble(self._problem(spec), [('generic', numpy.asarray(D)) for D in Ds], [('generic', numpy.asarray(C)) for C in Cs])
        return [m[0] for m in mats], list(vecs)
