'''Quadrature tables (host side, one-off).

Mirrors the reference's producers of integration points for tensor-product
elements: ``points.gauss1`` (src/nutils/points.py:342-355: ``degree//2+1``
Gauss-Legendre points mapped to [0,1]) and ``TensorPoints``
(points.py:144-164: C-order outer product).  The reference obtains the nodes
from a Golub-Welsch eigenproblem; here they come from numpy's Gauss-Legendre
routine, which agrees to rounding (checked in tests/test_tables.py).
'''

import functools
import numpy


@functools.lru_cache(maxsize=32)
def gauss1(degree):
    '''1-D Gauss rule on [0,1] exact for polynomials up to `degree`.

    Returns (points[nq], weights[nq]) with nq = degree//2 + 1, ascending.'''
    if degree < 0:
        raise ValueError('quadrature degree must be nonnegative')
    n = degree // 2 + 1
    x, w = numpy.polynomial.legendre.leggauss(n)
    x = (x + 1) * .5
    w = w * .5
    x.setflags(write=False)
    w.setflags(write=False)
    return x, w


def tensor_gauss(ndims, degree):
    'per-dimension (points, weights) of the tensor Gauss rule; degree may be an int or per-dim sequence'
    degrees = [degree] * ndims if numpy.ndim(degree) == 0 else list(degree)
    if len(degrees) != ndims:
        raise ValueError('degree does not match the number of dimensions')
    return [gauss1(int(d)) for d in degrees]


def tensor_points(rules):
    'flat C-order (coords[nq, ndims], weights[nq]) of a per-dim rule list (points.py:144-164)'
    pts = numpy.stack(numpy.meshgrid(*[r[0] for r in rules], indexing='ij'), axis=-1).reshape(-1, len(rules))
    wts = functools.reduce(numpy.multiply.outer, [r[1] for r in rules]).ravel()
    return pts, wts
