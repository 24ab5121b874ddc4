'''CPU restatement of the ``nutils_poly`` contract (numpy; TEST INFRASTRUCTURE ONLY).

The reference (evalf/nutils @ 37d1cc5) delegates every polynomial operation to
the third-party Rust/pyo3 package ``nutils-poly`` (``pyproject.toml:13``,
``nutils-poly >=1,<2``; no lock file, so no exact pin), whose source is NOT
under /root/reference and which is not installed in this image.  This module
restates its published behaviour from the reference's own documentation and
call sites:

* polynomial definition and coefficient order: ``evaluable.py:4328-4350``
  (``Polyval`` docstring) -- reverse lexicographic order, last variable most
  significant, highest power first;
* ``eval_outer(coeffs, points)``: ``evaluable.py:4365-4374``,
  ``element.py:203``;
* ``MulPlan`` / ``MulVar``: ``evaluable.py:4498-4563``, ``function.py:3098``;
* ``GradPlan``: ``evaluable.py:4584-4634``;
* ``ncoeffs`` / ``degree``: ``evaluable.py:4425-4495``;
* ``mul_different_vars`` / ``mul_same_vars`` / ``change_degree``:
  ``element.py:735``, ``topology.py:2202, 2503-2505, 3071-3072``.

Parity status: the reference's own tests compare these operations against
``nutils_poly`` itself (``tests/test_evaluable.py:588-599, 1334-1356``), so at
the raw-coefficient level parity is UNPINNED by the reference; it is pinned
end-to-end by the reference's known-answer tests that pass through it
(``tests/test_function.py:1574-1585``, ``tests/test_basis.py``,
``examples/laplace.py`` goldens) -- see tests/test_oracle_reference.py.

Nothing under nutils_b200/ imports this file.
'''

import enum
import functools
import itertools
import math
import numpy


def ncoeffs(nvars, degree):
    'number of coefficients of a polynomial in `nvars` variables of total degree `degree`'
    nvars = int(nvars)
    degree = int(degree)
    if nvars < 0 or degree < 0:
        raise ValueError('nvars and degree must be nonnegative')
    return math.comb(nvars + degree, nvars)


def degree(nvars, ncoeffs_):
    'inverse of ncoeffs; raises ValueError if `ncoeffs_` is not a valid count'
    nvars = int(nvars)
    n = int(ncoeffs_)
    if nvars == 0:
        if n == 1:
            return 0
        raise ValueError('invalid number of coefficients')
    p = 0
    while True:
        m = math.comb(nvars + p, nvars)
        if m == n:
            return p
        if m > n:
            raise ValueError('invalid number of coefficients')
        p += 1


@functools.lru_cache(maxsize=None)
def _powers(nvars, deg):
    '''integer powers per coefficient, shape (ncoeffs, nvars)

    Order: descending in (k_{n-1}, ..., k_0), i.e. last variable most
    significant and highest power first.
    '''
    if nvars == 0:
        return numpy.zeros((1, 0), dtype=int)
    pw = [k for k in itertools.product(range(deg + 1), repeat=nvars) if sum(k) <= deg]
    pw.sort(key=lambda k: k[::-1], reverse=True)
    a = numpy.array(pw, dtype=int).reshape(len(pw), nvars)
    a.setflags(write=False)
    return a


@functools.lru_cache(maxsize=None)
def _index(nvars, deg):
    return {tuple(k): i for i, k in enumerate(_powers(nvars, deg))}


def _vander(points, nvars, deg):
    'monomials at points: shape points.shape[:-1] + (ncoeffs,)'
    pw = _powers(nvars, deg)
    points = numpy.asarray(points, dtype=float)
    if points.shape[-1] != nvars:
        raise ValueError('points has wrong number of variables')
    # powers of each variable up to deg
    ladder = numpy.ones(points.shape + (deg + 1,), dtype=float)
    for k in range(1, deg + 1):
        ladder[..., k] = ladder[..., k - 1] * points
    out = numpy.ones(points.shape[:-1] + (len(pw),), dtype=float)
    for i in range(nvars):
        out *= ladder[..., i, pw[:, i]]
    return out


def eval_outer(coeffs, points):
    'evaluate: result shape points.shape[:-1] + coeffs.shape[:-1]'
    coeffs = numpy.asarray(coeffs)
    points = numpy.asarray(points, dtype=float)
    nvars = points.shape[-1]
    deg = degree(nvars, coeffs.shape[-1])
    V = _vander(points, nvars, deg)
    return numpy.tensordot(V, coeffs, axes=([-1], [-1]))


def eval(coeffs, points):
    'pointwise (broadcasting) evaluation'
    coeffs = numpy.asarray(coeffs)
    points = numpy.asarray(points, dtype=float)
    nvars = points.shape[-1]
    deg = degree(nvars, coeffs.shape[-1])
    V = _vander(points, nvars, deg)
    return (V * coeffs).sum(-1)


class MulVar(enum.Enum):
    Left = 0
    Right = 1
    Both = 2

    def __repr__(self):
        return 'MulVar.' + self.name

    def __getnewargs__(self):
        return (self.name,)


class MulPlan:
    '''Plan for the product of a polynomial in the variables marked Left/Both
    with one in the variables marked Right/Both.'''

    def __init__(self, vars, degree_left, degree_right):
        self.vars = tuple(vars)
        self.degree_left = dl = int(degree_left)
        self.degree_right = dr = int(degree_right)
        nvars = len(self.vars)
        lvars = [i for i, v in enumerate(self.vars) if v != MulVar.Right]
        rvars = [i for i, v in enumerate(self.vars) if v != MulVar.Left]
        pl = _powers(len(lvars), dl)
        pr = _powers(len(rvars), dr)
        full_l = numpy.zeros((len(pl), nvars), dtype=int)
        full_l[:, lvars] = pl
        full_r = numpy.zeros((len(pr), nvars), dtype=int)
        full_r[:, rvars] = pr
        idx = _index(nvars, dl + dr)
        self.nout = ncoeffs(nvars, dl + dr)
        self.nleft = len(pl)
        self.nright = len(pr)
        target = numpy.empty((len(pl), len(pr)), dtype=int)
        for i, kl in enumerate(full_l):
            for j, kr in enumerate(full_r):
                target[i, j] = idx[tuple(kl + kr)]
        # one-hot scatter matrix (nleft*nright, nout)
        S = numpy.zeros((len(pl) * len(pr), self.nout), dtype=float)
        S[numpy.arange(S.shape[0]), target.ravel()] = 1.
        self._S = S

    def __call__(self, coeffs_left, coeffs_right):
        cl = numpy.asarray(coeffs_left)
        cr = numpy.asarray(coeffs_right)
        if cl.shape[-1] != self.nleft or cr.shape[-1] != self.nright:
            raise ValueError('coefficient arrays do not match the plan')
        prod = cl[..., :, None] * cr[..., None, :]
        return prod.reshape(prod.shape[:-2] + (-1,)) @ self._S


@functools.lru_cache(maxsize=None)
def _mulplan(vars, dl, dr):
    return MulPlan(vars, dl, dr)


def mul(coeffs_left, coeffs_right, vars):
    vars = tuple(vars)
    nl = sum(v != MulVar.Right for v in vars)
    nr = sum(v != MulVar.Left for v in vars)
    dl = degree(nl, numpy.shape(coeffs_left)[-1])
    dr = degree(nr, numpy.shape(coeffs_right)[-1])
    return _mulplan(vars, dl, dr)(coeffs_left, coeffs_right)


def mul_different_vars(coeffs_left, coeffs_right, nleft, nright):
    return mul(coeffs_left, coeffs_right, (MulVar.Left,) * nleft + (MulVar.Right,) * nright)


def mul_same_vars(coeffs_left, coeffs_right, nvars):
    return mul(coeffs_left, coeffs_right, (MulVar.Both,) * nvars)


class GradPlan:
    'Plan for the gradient: c[..., ncoeffs] -> [..., nvars, ncoeffs(nvars, max(degree-1, 0))]'

    def __init__(self, nvars, degree):
        self.nvars = nvars = int(nvars)
        self.degree = deg = int(degree)
        self.nin = ncoeffs(nvars, deg)
        dout = max(deg - 1, 0)
        self.nout = ncoeffs(nvars, dout)
        G = numpy.zeros((nvars, self.nin, self.nout), dtype=float)
        if deg > 0:
            idx = _index(nvars, dout)
            for i, k in enumerate(_powers(nvars, deg)):
                for v in range(nvars):
                    if k[v] > 0:
                        kk = k.copy()
                        kk[v] -= 1
                        G[v, i, idx[tuple(kk)]] = k[v]
        self._G = G

    def __call__(self, coeffs):
        c = numpy.asarray(coeffs)
        if c.shape[-1] != self.nin:
            raise ValueError('coefficient array does not match the plan')
        return numpy.einsum('...i,vio->...vo', c, self._G)


@functools.lru_cache(maxsize=None)
def _gradplan(nvars, deg):
    return GradPlan(nvars, deg)


def grad(coeffs, nvars):
    deg = degree(nvars, numpy.shape(coeffs)[-1])
    return _gradplan(nvars, deg)(coeffs)


def change_degree(coeffs, nvars, newdegree):
    'embed coefficients in the (larger) coefficient space of `newdegree`'
    c = numpy.asarray(coeffs)
    deg = degree(nvars, c.shape[-1])
    if newdegree < deg:
        raise ValueError('cannot lower the degree')
    idx = _index(nvars, newdegree)
    out = numpy.zeros(c.shape[:-1] + (ncoeffs(nvars, newdegree),), dtype=c.dtype)
    out[..., [idx[tuple(k)] for k in _powers(nvars, deg)]] = c
    return out


def composition_with_inner_matrix(inner_coeffs, inner_nvars, outer_nvars, degree_):
    '''Matrix M such that M @ c are the coefficients of p(A [x; 1]) given those of p.

    ``inner_coeffs`` has shape (outer_nvars, ncoeffs(inner_nvars, 1)): row v are the
    degree-1 coefficients (in this module's order) of outer variable v as a
    function of the inner variables (``transform.py:185``).
    '''
    inner = numpy.asarray(inner_coeffs, dtype=float)
    nin = ncoeffs(inner_nvars, degree_)
    pw = _powers(outer_nvars, degree_)
    M = numpy.zeros((nin, len(pw)), dtype=float)
    one = numpy.zeros(ncoeffs(inner_nvars, 0))
    one[0] = 1.
    for col, k in enumerate(pw):
        acc, accdeg = one, 0
        for v in range(outer_nvars):
            for _ in range(k[v]):
                acc = _mulplan((MulVar.Both,) * inner_nvars, accdeg, 1)(acc, inner[v])
                accdeg += 1
        M[:, col] = change_degree(acc, inner_nvars, degree_)
    return M
