'''Mesh generators of the accelerated path (reference: src/nutils/mesh.py:34-60).'''

import numpy
from . import topology, function


def rectilinear(richshape, periodic=(), space='X'):
    '''Rectilinear mesh: returns (topology, geometry) like ``nutils.mesh.rectilinear``.

    Every entry of `richshape` is an element count (vertices 0..n) or a vertex
    array.  The geometry is the multilinear interpolation of the vertex grid --
    the reference builds exactly this for non-integer vertices (a degree-1
    spline times nodal coordinates, mesh.py:55-57); integer grids are the same
    function.  ``geom.nodes`` is read-only; a deformed mesh is a new geometry,
    ``topo.nodal_geometry(geom.nodes + displacement)``.'''
    if periodic:
        raise NotImplementedError('periodic meshes are outside the accelerated path')
    verts = [numpy.arange(v + 1, dtype=float) if numpy.ndim(v) == 0 else numpy.asarray(v, dtype=float) for v in richshape]
    if not 1 <= len(verts) <= 3:
        raise NotImplementedError('1 to 3 dimensions')
    topo = topology.StructuredTopology(tuple(len(v) - 1 for v in verts), space=space)
    nodes = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    return topo, function.Geometry(topo, nodes)


def line(nodes, space='X'):
    topo, geom = rectilinear([nodes], space=space)
    return topo, geom


def unitsquare(nelems, etype='square'):
    if etype != 'square':
        raise NotImplementedError('only structured square elements are on the accelerated path')
    return rectilinear([numpy.linspace(0, 1, nelems + 1)] * 2)
