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


def _all_golden_names():
    return sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(GOLDEN, '*.npz')))


def _kind(name):
    with numpy.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as f:
        return str(f['kind'])


def golden_names():
    'structured cases: every element of the grid, one tensor rule, multilinear geometry'
    return [n for n in _all_golden_names() if _kind(n) in ('scalar', 'elasticity')]


def elemset_golden_names():
    'element-set cases: trimmed topologies with ragged points and pruned numbering, NURBS'
    return [n for n in _all_golden_names() if _kind(n).startswith('elemset') and _kind(n) != 'elemset_eval']


def eval_golden_names():
    'Sample.eval cases: coordinates, weights, field values and gradients at the points'
    return [n for n in _all_golden_names() if _kind(n) == 'elemset_eval']


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


def elemset_problem_from_golden(g):
    'oracle Problem of an element-set golden (kind elemset_scalar / elemset_elasticity)'
    nelems = tuple(int(n) for n in g['nelems'])
    degree = int(g['degree'])
    ndims = len(nelems)
    b1 = bases_1d(nelems, degree, str(g['btype']))
    rules = points.tensor_gauss(ndims, int(g['qdegree']))
    kw = dict(ncomp=int(g['ncomp']))
    for key in 'elem_ids', 'qoff', 'qcoords', 'qweights', 'renumber', 'scale', 'face_dim', 'normals':
        if key in g:
            kw[key] = g[key]
    if 'renumber' in g:
        kw['nbasis_new'] = int(g['nbasis_new'])
    if 'rational' in g:
        kw['rational'] = int(g['rational'])
    nodes = g['nodes'] if 'nodes' in g else None
    if 'gctrl' in g:
        gb = bases_1d(nelems, int(g['gdegree']), 'spline')
        kw['geom_spline'] = dict(degree=[b.degree for b in gb], coeffs=[b.coeffs for b in gb], setidx=[b.setidx for b in gb], start=[b.start for b in gb],
                                 ndofs_d=[b.ndofs for b in gb], ctrl=g['gctrl'], weights=g['gweights'] if 'gweights' in g else None)
        nodes = numpy.zeros((ndims,) + (0,) * ndims)
    return fem_oracle.Problem(nelems, [degree] * ndims, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], nodes, **kw)


def elemset_forms(g, ndims):
    '''(matrix coefficient tensors D, vector coefficient tensors C, expected values, expected rhs) of an element-set golden;
    elasticity goldens hold the jacobian/residual of int (grad_j(v_i) sigma_ij - v_i q_i) dV at u=0'''
    from nutils_b200 import engine
    if str(g['kind']) == 'elemset_scalar':
        return [engine.form_stiffness(ndims), engine.form_mass(ndims)], [engine.form_load(ndims)], [g['K_values'], g['M_values']], g['F']
    if str(g['kind']) == 'elemset_boundary':   # int_G g N_i N_j dS, int_G g N_i dS
        return [engine.form_mass(ndims)], [engine.form_load(ndims)], [g['M_values']], g['F']
    if str(g['kind']) == 'elemset_varcoef':    # int g grad N_i . grad N_j dV, int g N_i dV
        return [engine.form_stiffness(ndims)], [engine.form_load(ndims)], [g['K_values']], g['F']
    C = numpy.zeros((ndims, ndims + 1))
    C[:, 0] = -numpy.asarray(g['load'])
    return [engine.form_elasticity(ndims, float(g['lmbda']), float(g['mu']))], [C], [g['K_values']], g['F']


def elemset_coefs(g, nmat, nvec):
    'pointwise coefficients per matrix / vector form of an element-set golden (None: constant coefficient tensors only)'
    if 'coef' not in g:
        return None, None
    return [g['coef']] * nmat, [g['coef']] * nvec


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
