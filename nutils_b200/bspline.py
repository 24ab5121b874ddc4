'''1-D B-spline tables (host side, one-off): the per-dimension inputs of the CUDA path.

What the reference computes in ``StructuredTopology.basis_spline``
(src/nutils/topology.py:2209-2324) and ``_localsplinebasis`` (:2327-2361) and
stores in ``function.StructuredBasis`` (function.py:3051-3058): per dimension a
list of local polynomial coefficient sets (one per element, deduplicated), the
first dof of every element and the number of dofs.  The local polynomials are
expressed in the element coordinate xi in [0,1] with the HIGHEST power first,
the same convention as the reference.

The implementation here is an independent Cox-de Boor recursion on ascending
coefficient arrays; tests/test_tables.py checks it against the reference's
tables when /root/reference is available and against partition of unity
otherwise.
'''

import numpy
from numpy.polynomial import polynomial as P


def local_polynomials(lknots):
    '''The p+1 B-splines that are nonzero on the knot span [lknots[p-1], lknots[p]].

    `lknots` holds the 2p knots around the span.  Returns an array (p+1, p+1):
    row a = coefficients of N_a(xi), highest power first, xi the local
    coordinate of the span.'''
    lknots = numpy.asarray(lknots, dtype=float)
    p, rem = divmod(len(lknots), 2)
    if rem:
        raise ValueError('local knot vector must have even length')
    if p == 0:
        return numpy.ones((1, 1))
    a, b = lknots[p - 1], lknots[p]
    if not b > a:
        raise ValueError('element size should be positive')
    if (numpy.diff(lknots) < -numpy.spacing(1)).any():
        raise ValueError('local knot vector should be non-decreasing')
    x = numpy.array([a, b - a])  # x(xi) = a + (b-a) xi, ascending coefficients
    N = [numpy.array([1.])]
    for j in range(1, p + 1):
        # raise degree j-1 -> j (de Boor): N_r <- saved + right_{r+1} temp, saved <- left_{j-r} temp
        new = []
        saved = numpy.zeros(1)
        for r in range(j):
            kr = lknots[p + r]        # right knot  U[i+r+1]
            kl = lknots[p - j + r]    # left knot   U[i+1-(j-r)]
            temp = N[r] / (kr - kl)
            right = P.polysub([kr], x)
            left = P.polysub(x, [kl])
            new.append(P.polyadd(saved, P.polymul(right, temp)))
            saved = P.polymul(left, temp)
        new.append(saved)
        N = new
    out = numpy.zeros((p + 1, p + 1))
    for r, c in enumerate(N):
        c = numpy.asarray(c)
        out[r, p + 1 - len(c):] = c[::-1]
    return out


class Basis1D:
    '''Spline space along one dimension of a structured topology.

    Attributes
    ----------
    degree, nelems, ndofs : int
    coeffs : float64[nsets, p+1, p+1]   unique local coefficient sets (highest power first)
    setidx : int32[nelems]               coefficient set of every element
    start  : int64[nelems]               first dof of every element (dofs are start..start+p, modulo ndofs if periodic)
    periodic : bool
    '''

    def __init__(self, degree, nelems, coeffs, setidx, start, ndofs, periodic=False):
        self.degree = int(degree)
        self.nelems = int(nelems)
        self.coeffs = numpy.ascontiguousarray(coeffs, dtype=float)
        self.setidx = numpy.ascontiguousarray(setidx, dtype=numpy.int32)
        self.start = numpy.ascontiguousarray(start, dtype=numpy.int64)
        self.ndofs = int(ndofs)
        self.periodic = bool(periodic)
        assert self.coeffs.shape[1:] == (self.degree + 1,) * 2
        assert self.setidx.shape == self.start.shape == (self.nelems,)

    def tabulate(self, points):
        '''values and xi-derivatives of every coefficient set at 1-D `points`: float64[nsets, 2, p+1, nq]'''
        points = numpy.asarray(points, dtype=float)
        p = self.degree
        tab = numpy.empty((len(self.coeffs), 2, p + 1, len(points)))
        dscale = numpy.arange(p, 0, -1.)
        for s, c in enumerate(self.coeffs):
            for a in range(p + 1):
                tab[s, 0, a] = numpy.polyval(c[a], points)
                tab[s, 1, a] = numpy.polyval(c[a, :-1] * dscale, points) if p else 0.
        return tab


def spline_basis_1d(nelems, degree, continuity=-1, knotvalues=None, knotmultiplicities=None, periodic=False):
    '''Tables for a 1-D spline space; argument meaning as in the reference
    (topology.py:2209-2324): `continuity` negative counts from the degree,
    default knots are uniform, default multiplicity is degree-continuity with
    open ends.'''
    p = int(degree)
    n = int(nelems)
    if p < 0 or n < 1:
        raise ValueError('invalid degree or element count')
    c = continuity + p if continuity < 0 else continuity
    if not -1 <= c < max(p, 1) and p > 0:
        raise ValueError('continuity out of range')
    k = numpy.arange(n + 1, dtype=float) if knotvalues is None else numpy.array(knotvalues, dtype=float)
    if len(k) != n + 1:
        raise ValueError('knot values do not match the topology size')
    if knotmultiplicities is None:
        m = numpy.repeat(p - c, n + 1)
    else:
        m = numpy.array(knotmultiplicities, dtype=int)
        if len(m) != n + 1 or m.min() <= 0 or m.max() > p + 1:
            raise ValueError('incorrect knot multiplicities')
    if p == 0:
        return Basis1D(0, n, numpy.ones((1, 1, 1)), numpy.zeros(n, dtype=numpy.int32), numpy.arange(n), n, False)
    if periodic:
        raise NotImplementedError('periodic spline spaces are outside the accelerated path')
    # open knot vector: the end knots are repeated p + 1 times in all, interior knot i mult[i] times
    mult = m.copy()
    mult[0] = mult[-1] = p
    nd = int(mult[:n].sum()) + 1
    km = numpy.repeat(k, mult).astype(float)
    m = mult
    offsets = numpy.cumsum(m[:n]) - m[0]
    sets = []
    keys = {}
    setidx = numpy.empty(n, dtype=numpy.int32)
    for ielem, offset in enumerate(offsets):
        lk = km[offset:offset + 2 * p]
        key = tuple(numpy.round((lk[1:-1] - lk[0]) / (lk[-1] - lk[0]) * 2**31).astype(numpy.int64))
        if key not in keys:
            keys[key] = len(sets)
            sets.append(local_polynomials(lk))
        setidx[ielem] = keys[key]
    return Basis1D(p, n, numpy.array(sets), setidx, offsets, nd, False)
