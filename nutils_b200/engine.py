'''Thin object layer over the C ABI: contexts, device-resident tables, assembly calls.

This is the host-side glue between the Nutils-style API (topology.py, sample.py,
function.py) and libb200fem.  It owns no numerics: tables are produced by
bspline.py / points.py, the arithmetic happens in the CUDA kernels.
'''

import ctypes
import weakref
import numpy

from . import _lib
from ._lib import check, as_f64, c_vp, c_i64


class Context:
    '''One CUDA context wrapper per process and device (b2_ctx).'''

    _instances = {}

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = c_vp()
        check(self.lib.b2_ctx_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)
        self._finalizer = weakref.finalize(self, self.lib.b2_ctx_destroy, h)

    @classmethod
    def get(cls, device=0):
        'shared context of a device'
        ctx = cls._instances.get(device)
        if ctx is None:
            ctx = cls._instances[device] = cls(device)
        return ctx

    def check(self, status):
        check(status, self.handle)

    def synchronize(self):
        self.check(self.lib.b2_ctx_synchronize(self.handle))

    def set_stream(self, cuda_stream):
        'run on an external stream (int handle, e.g. torch.cuda.current_stream().cuda_stream); None restores the own stream'
        if cuda_stream is None:
            handle = 0           # the context's own stream
        elif int(cuda_stream) == 0:
            handle = 1           # cudaStreamLegacy: the default stream torch hands out as 0
        else:
            handle = int(cuda_stream)
        self.check(self.lib.b2_ctx_set_stream(self.handle, c_vp(handle)))

    def set_option(self, name, value):
        self.check(self.lib.b2_ctx_set_option(self.handle, name.encode(), int(value)))

    def timer_start(self):
        self.check(self.lib.b2_ctx_timer_start(self.handle))

    def timer_stop(self):
        ms = ctypes.c_float()
        self.check(self.lib.b2_ctx_timer_stop(self.handle, ctypes.byref(ms)))
        return float(ms.value)

    def kernel_time(self):
        '(summed device milliseconds, launches) of the assembly kernels since the last call; needs set_option("time_kernels", 1)'
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        self.check(self.lib.b2_ctx_kernel_time(self.handle, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def launch_count(self):
        return int(self.lib.b2_ctx_launch_count(self.handle))

    def flush_l2(self):
        self.check(self.lib.b2_flush_l2(self.handle))

    # raw memory -----------------------------------------------------------------------------------

    def host_empty(self, shape, dtype=numpy.float64):
        'pinned host array (freed when garbage collected)'
        dtype = numpy.dtype(dtype)
        n = int(numpy.prod(shape)) * dtype.itemsize
        p = c_vp()
        self.check(self.lib.b2_host_alloc(self.handle, n, ctypes.byref(p)))
        buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
        arr = numpy.frombuffer(buf, dtype=dtype, count=int(numpy.prod(shape))).reshape(shape)
        weakref.finalize(buf, self.lib.b2_host_free, self.handle, p)
        return arr

    def device_alloc(self, nbytes):
        p = c_vp()
        self.check(self.lib.b2_device_alloc(self.handle, int(nbytes), ctypes.byref(p)))
        return DeviceBuffer(self, p, int(nbytes))


class DeviceBuffer:
    'device memory owned through the C ABI (used when torch is not wanted)'

    def __init__(self, ctx, ptr, nbytes):
        self.ctx = ctx
        self.ptr = ptr
        self.nbytes = nbytes
        self._finalizer = weakref.finalize(self, ctx.lib.b2_device_free, ctx.handle, ptr)

    def zero(self):
        self.ctx.check(self.ctx.lib.b2_memset_zero(self.ctx.handle, self.ptr, self.nbytes))

    def to_host(self, dtype=numpy.float64, out=None):
        dtype = numpy.dtype(dtype)
        if out is None:
            out = numpy.empty(self.nbytes // dtype.itemsize, dtype=dtype)
        self.ctx.check(self.ctx.lib.b2_memcpy_d2h(self.ctx.handle, out.ctypes.data_as(c_vp), self.ptr, self.nbytes))
        return out

    def from_host(self, arr):
        arr = numpy.ascontiguousarray(arr)
        assert arr.nbytes == self.nbytes
        self.ctx.check(self.ctx.lib.b2_memcpy_h2d(self.ctx.handle, self.ptr, arr.ctypes.data_as(c_vp), self.nbytes))


def _devptr(x):
    'device pointer of a DeviceBuffer, a torch tensor or an int'
    if isinstance(x, DeviceBuffer):
        return x.ptr
    if hasattr(x, 'data_ptr'):
        return c_vp(x.data_ptr())
    return c_vp(int(x))


def _create_basis(ctx, bases, ncomp):
    'b2_basis handle of a tensor-product space given as per-dimension bspline.Basis1D tables'
    if any(b.periodic for b in bases):
        raise _lib.B200Error('unsupported configuration: periodic bases')
    nd = len(bases)
    nelems = numpy.array([b.nelems for b in bases], dtype=numpy.int64)
    degree = numpy.array([b.degree for b in bases], dtype=numpy.int32)
    nsets = numpy.array([len(b.coeffs) for b in bases], dtype=numpy.int32)
    ndofs = numpy.array([b.ndofs for b in bases], dtype=numpy.int64)
    coeffs = [as_f64(b.coeffs) for b in bases]
    setidx = [numpy.ascontiguousarray(b.setidx, dtype=numpy.int32) for b in bases]
    start = [numpy.ascontiguousarray(b.start, dtype=numpy.int64) for b in bases]
    h = c_vp()
    ctx.check(ctx.lib.b2_basis_create(ctx.handle, nd, nelems.ctypes.data_as(_lib.p_i64), degree.ctypes.data_as(_lib.p_i32), nsets.ctypes.data_as(_lib.p_i32),
                                      _lib.ptr_array(coeffs, _lib.p_f64), _lib.ptr_array(setidx, _lib.p_i32), _lib.ptr_array(start, _lib.p_i64),
                                      ndofs.ctypes.data_as(_lib.p_i64), int(ncomp), ctypes.byref(h)))
    return h


class Plan:
    '''Device-resident description of one structured assembly problem:
    spline space (b2_basis) + tensor quadrature (b2_quad) + nodal geometry (b2_geom)
    + analytic CSR pattern (b2_pattern).

    bases : sequence of bspline.Basis1D, one per dimension
    rules : sequence of (points, weights), one per dimension
    nodes : float64[ndims, n0+1, n1+1, ...]
    '''

    def __init__(self, ctx, bases, rules, nodes, ncomp=1):
        self.ctx = ctx
        lib = ctx.lib
        self.ndims = nd = len(bases)
        self.ncomp = int(ncomp)
        self.bases = list(bases)
        self.rules = [(as_f64(x), as_f64(w)) for x, w in rules]
        self.nelems = tuple(b.nelems for b in bases)
        self.ntotal = int(numpy.prod(self.nelems))
        nodes = as_f64(nodes)
        if nodes.shape != (nd,) + tuple(n + 1 for n in self.nelems):
            raise ValueError('nodes must have shape (ndims, nelems_0+1, ...), got {}'.format(nodes.shape))
        self.nodes = nodes
        nelems = numpy.array(self.nelems, dtype=numpy.int64)
        self.basis = h = _create_basis(ctx, bases, self.ncomp)
        self._fin = [weakref.finalize(self, lib.b2_basis_destroy, h)]
        nq = numpy.array([len(x) for x, w in self.rules], dtype=numpy.int32)
        h = c_vp()
        ctx.check(lib.b2_quad_create_tensor(ctx.handle, nd, nq.ctypes.data_as(_lib.p_i32), _lib.ptr_array([x for x, w in self.rules], _lib.p_f64),
                                            _lib.ptr_array([w for x, w in self.rules], _lib.p_f64), ctypes.byref(h)))
        self.quad = h
        self._fin.append(weakref.finalize(self, lib.b2_quad_destroy, h))
        h = c_vp()
        ctx.check(lib.b2_geom_create_nodal(ctx.handle, nd, nelems.ctypes.data_as(_lib.p_i64), nodes.ctypes.data_as(c_vp), ctypes.byref(h)))
        self.geom = h
        self._fin.append(weakref.finalize(self, lib.b2_geom_destroy, h))
        h = c_vp()
        ctx.check(lib.b2_pattern_create(ctx.handle, self.basis, ctypes.byref(h)))
        self.pattern = h
        self._fin.append(weakref.finalize(self, lib.b2_pattern_destroy, h))
        self.nnz = int(lib.b2_pattern_nnz(h))
        self.ndofs = int(lib.b2_pattern_nrows(h))
        self._csr = None

    # pattern --------------------------------------------------------------------------------------

    def csr_pattern(self):
        '(rowptr int64[ndofs+1], colidx int64[nnz]) on the host, cached'
        if self._csr is None:
            rowptr = numpy.empty(self.ndofs + 1, dtype=numpy.int64)
            colidx = numpy.empty(self.nnz, dtype=numpy.int64)
            self.ctx.check(self.ctx.lib.b2_pattern_export_host(self.pattern, rowptr.ctypes.data_as(c_vp), colidx.ctypes.data_as(c_vp)))
            self._csr = rowptr, colidx
        return self._csr

    def csr_pattern_device(self, rowptr_dev, colidx_dev):
        self.ctx.check(self.ctx.lib.b2_pattern_export_device(self.pattern, _devptr(rowptr_dev), _devptr(colidx_dev)))

    def row_offset(self, row):
        'rowptr[row], computed analytically on the host'
        off = int(self.ctx.lib.b2_pattern_row_offset(self.pattern, int(row)))
        if off < 0:
            raise ValueError('row out of range')
        return off

    # device-resident matrix operations (b2_spmv_device, b2_diagonal_device, b2_cg_device) ------------

    def spmv_device(self, values, x, y):
        'y = A x for device arrays (DeviceBuffer / torch tensor / pointer); asynchronous on the context stream'
        self.ctx.check(self.ctx.lib.b2_spmv_device(self.ctx.handle, self.pattern, _devptr(values), _devptr(x), _devptr(y)))

    def diagonal_device(self, values, diag):
        self.ctx.check(self.ctx.lib.b2_diagonal_device(self.ctx.handle, self.pattern, _devptr(values), _devptr(diag)))

    def cg_device(self, values, rhs, x, constrained=None, atol=0., rtol=0., maxiter=0):
        '''Jacobi-preconditioned CG on the device: x holds lhs0 (constrained values included) on entry, the solution on return.
        Returns (iterations, residual norm); raises ToleranceNotReached if an explicit tolerance is not met.'''
        it = ctypes.c_int()
        res = ctypes.c_double()
        status = self.ctx.lib.b2_cg_device(self.ctx.handle, self.pattern, _devptr(values), None if rhs is None else _devptr(rhs), _devptr(x),
                                           None if constrained is None else _devptr(constrained), float(atol), float(rtol), int(maxiter),
                                           ctypes.byref(it), ctypes.byref(res))
        self.last_cg = int(it.value), float(res.value)
        self.ctx.check(status)
        return self.last_cg

    def update_nodes(self, nodes):
        nodes = as_f64(nodes)
        assert nodes.shape == self.nodes.shape
        self.nodes = nodes
        self.ctx.check(self.ctx.lib.b2_geom_update_nodal(self.geom, nodes.ctypes.data_as(c_vp)))

    # assembly -------------------------------------------------------------------------------------

    def _form_args(self, Ds, Cs):
        na = self.ndims + 1
        nc = self.ncomp
        Ds = [as_f64(D).reshape(nc, na, nc, na) for D in Ds]
        Cs = [as_f64(C).reshape(nc, na) for C in Cs]
        return Ds, Cs, _lib.ptr_array(Ds, _lib.p_f64), _lib.ptr_array(Cs, _lib.p_f64)

    def assemble_host(self, Ds=(), Cs=(), elem_range=None, out_values=None, out_rhs=None):
        '''Assemble into host arrays through b2_assemble_host: returns ([values...], [rhs...]).'''
        Ds, Cs, pD, pC = self._form_args(Ds, Cs)
        vals = out_values if out_values is not None else [numpy.empty(self.nnz) for _ in Ds]
        rhs = out_rhs if out_rhs is not None else [numpy.empty(self.ndofs) for _ in Cs]
        e0, e1 = elem_range if elem_range is not None else (0, -1)
        pv = (c_vp * max(len(vals), 1))(*[v.ctypes.data_as(c_vp) for v in vals])
        pr = (c_vp * max(len(rhs), 1))(*[r.ctypes.data_as(c_vp) for r in rhs])
        self.ctx.check(self.ctx.lib.b2_assemble_host(self.ctx.handle, self.pattern, self.basis, self.quad, self.geom, c_i64(e0), c_i64(e1),
                                                     len(Ds), pD, pv, len(Cs), pC, pr))
        return vals, rhs

    def assemble_device(self, Ds=(), Cs=(), values=(), rhs=(), elem_range=None):
        '''Accumulate into device arrays (DeviceBuffer / torch tensors / raw pointers); asynchronous on the context stream.'''
        Ds, Cs, pD, pC = self._form_args(Ds, Cs)
        assert len(values) == len(Ds) and len(rhs) == len(Cs)
        e0, e1 = elem_range if elem_range is not None else (0, -1)
        pv = (c_vp * max(len(values), 1))(*[_devptr(v) for v in values])
        pr = (c_vp * max(len(rhs), 1))(*[_devptr(r) for r in rhs])
        self.ctx.check(self.ctx.lib.b2_assemble_device(self.ctx.handle, self.pattern, self.basis, self.quad, self.geom, c_i64(e0), c_i64(e1),
                                                       len(Ds), pD, pv, len(Cs), pC, pr))

    def assemble_rows_device(self, Ds=(), Cs=(), values=(), rhs=(), plane_range=None):
        '''Owner-computes assembly (b2_assemble_rows_device): WRITES every stored value of the dof rows whose index along
        dimension 0 lies in plane_range (default: all); no zero-fill needed, rows outside the range are untouched.'''
        Ds, Cs, pD, pC = self._form_args(Ds, Cs)
        assert len(values) == len(Ds) and len(rhs) == len(Cs)
        p0, p1 = plane_range if plane_range is not None else (0, -1)
        pv = (c_vp * max(len(values), 1))(*[_devptr(v) for v in values])
        pr = (c_vp * max(len(rhs), 1))(*[_devptr(r) for r in rhs])
        self.ctx.check(self.ctx.lib.b2_assemble_rows_device(self.ctx.handle, self.pattern, self.basis, self.quad, self.geom, c_i64(p0), c_i64(p1),
                                                            len(Ds), pD, pv, len(Cs), pC, pr))


class ElemSetPlan:
    '''Device-resident description of an assembly problem on an ELEMENT SET of a structured grid (b2_elemset):
    a subset of the elements (trimmed topologies), per-element point sets of different length (cut cells),
    the pruned dof numbering, rational (NURBS) scaling, and a nodal or (rational) spline geometry.

    bases     : sequence of bspline.Basis1D of the PARENT tensor space
    nodes     : float64[ndims, n0+1, ...] nodal geometry, or None when `geom_spline` is given
    elem_ids  : int64[nsel] selected elements (C-order indices, increasing); None = all
    qoff, qcoords, qweights : ragged points in element-local coordinates; None = tensor `rules`
    renumber  : int64[nbasis_parent] -> new index (negative or >= nbasis_new: dropped); None = identity
    scale     : float64[nbasis_parent] numerator weights c_i; rational: 0 none, 1 divide by sum_j c_j B_j,
                2 divide by the weight function of the rational spline geometry
    geom_spline : (bases_g, ctrl[ndims, nbasis_g], weights[nbasis_g] or None)
    '''

    def __init__(self, ctx, bases, nodes=None, ncomp=1, rules=None, elem_ids=None, qoff=None, qcoords=None, qweights=None,
                 renumber=None, nbasis_new=None, scale=None, rational=0, geom_spline=None):
        self.ctx = ctx
        lib = ctx.lib
        self.ndims = nd = len(bases)
        self.ncomp = int(ncomp)
        self.nelems = tuple(b.nelems for b in bases)
        self._fin = []
        self._keep = []
        self.basis = self._make_basis(bases, self.ncomp)
        nelems = numpy.array(self.nelems, dtype=numpy.int64)
        # quadrature
        self.quad = None
        if qoff is None:
            if rules is None:
                raise ValueError('either ragged points (qoff, qcoords, qweights) or tensor rules are needed')
            rules = [(as_f64(x), as_f64(w)) for x, w in rules]
            self.rules = rules
            nq = numpy.array([len(x) for x, w in rules], dtype=numpy.int32)
            h = c_vp()
            ctx.check(lib.b2_quad_create_tensor(ctx.handle, nd, nq.ctypes.data_as(_lib.p_i32), _lib.ptr_array([x for x, w in rules], _lib.p_f64),
                                                _lib.ptr_array([w for x, w in rules], _lib.p_f64), ctypes.byref(h)))
            self.quad = h
            self._fin.append(weakref.finalize(self, lib.b2_quad_destroy, h))
        # geometry
        h = c_vp()
        if geom_spline is not None:
            gbases, ctrl, gw = geom_spline
            self.gbasis = self._make_basis(gbases, 1)
            ctrl = as_f64(ctrl)
            nbg = int(numpy.prod([b.ndofs for b in gbases]))
            if ctrl.shape != (nd, nbg):
                raise ValueError('control points must have shape (ndims, nbasis_g)')
            gw = None if gw is None else as_f64(gw)
            ctx.check(lib.b2_geom_create_spline(ctx.handle, self.gbasis, ctrl.ctypes.data_as(c_vp), None if gw is None else gw.ctypes.data_as(c_vp), ctypes.byref(h)))
        else:
            nodes = as_f64(nodes)
            if nodes.shape != (nd,) + tuple(n + 1 for n in self.nelems):
                raise ValueError('nodes must have shape (ndims, nelems_0+1, ...), got {}'.format(nodes.shape))
            ctx.check(lib.b2_geom_create_nodal(ctx.handle, nd, nelems.ctypes.data_as(_lib.p_i64), nodes.ctypes.data_as(c_vp), ctypes.byref(h)))
        self.geom = h
        self._fin.append(weakref.finalize(self, lib.b2_geom_destroy, h))
        # element set
        ntot = int(numpy.prod(self.nelems))
        i64 = lambda a: None if a is None else numpy.ascontiguousarray(a, dtype=numpy.int64)
        f64 = lambda a: None if a is None else as_f64(a)
        elem_ids, qoff, renumber = i64(elem_ids), i64(qoff), i64(renumber)
        qcoords, qweights, scale = f64(qcoords), f64(qweights), f64(scale)
        nsel = ntot if elem_ids is None else len(elem_ids)
        nbasis = int(numpy.prod([b.ndofs for b in bases]))
        if qoff is not None and (len(qoff) != nsel + 1 or qcoords.shape != (int(qoff[-1]), nd) or qweights.shape != (int(qoff[-1]),)):
            raise ValueError('ragged quadrature arrays do not match the element selection')
        if renumber is not None and len(renumber) != nbasis:
            raise ValueError('renumber must have one entry per parent basis function')
        if scale is not None and len(scale) != nbasis:
            raise ValueError('scale must have one entry per parent basis function')
        if renumber is not None and nbasis_new is None:
            nbasis_new = int(renumber[(renumber >= 0) & (renumber < nbasis)].max()) + 1 if len(renumber) else 0
        ptr = lambda a: None if a is None else a.ctypes.data_as(c_vp)
        h = c_vp()
        ctx.check(lib.b2_elemset_create(ctx.handle, self.basis, nsel, ptr(elem_ids), ptr(qoff), ptr(qcoords), ptr(qweights), ptr(renumber),
                                        int(nbasis_new or 0), ptr(scale), int(rational), ctypes.byref(h)))
        self.elemset = h
        self._fin.append(weakref.finalize(self, lib.b2_elemset_destroy, h))
        self.nsel = nsel
        self.npoints = int(lib.b2_elemset_npoints(h))
        h = c_vp()
        ctx.check(lib.b2_pattern_create_elemset(ctx.handle, self.elemset, ctypes.byref(h)))
        self.pattern = h
        self._fin.append(weakref.finalize(self, lib.b2_pattern_destroy, h))
        self.nnz = int(lib.b2_pattern_nnz(h))
        self.ndofs = int(lib.b2_pattern_nrows(h))
        self._csr = None

    def set_faces(self, face_dim):
        '''boundary integrals: face_dim[k] = -1 (volume) or the reference direction normal to the face the points of selected
        element k lie on; their weights are multiplied by the surface measure (b2_elemset_set_faces).  None resets.'''
        fd = None if face_dim is None else numpy.ascontiguousarray(face_dim, dtype=numpy.int8)
        if fd is not None and fd.shape != (self.nsel,):
            raise ValueError('one face code per selected element is needed')
        self.ctx.check(self.ctx.lib.b2_elemset_set_faces(self.elemset, None if fd is None else fd.ctypes.data_as(c_vp)))

    def set_normals(self, normals):
        '''immersed boundaries: per point the reference-space normal of its facet scaled by the facet's reference measure
        (float64[npoints, ndims]); the weights are multiplied by |det J| |J^-T n| (b2_elemset_set_normals).  None resets.'''
        if normals is None:
            self.ctx.check(self.ctx.lib.b2_elemset_set_normals(self.elemset, None, 0))
            return
        nr = as_f64(normals)
        if nr.shape != (self.npoints, self.ndims):
            raise ValueError('one normal per quadrature point is needed')
        self.ctx.check(self.ctx.lib.b2_elemset_set_normals(self.elemset, nr.ctypes.data_as(c_vp), len(nr)))

    def set_coefficient(self, kind, index, coef):
        '''a scalar per quadrature point multiplying matrix form `index` (kind 'matrix') or vector form `index` (kind 'vector') of
        the following assembly calls (b2_elemset_set_coefficient); coef None removes it.'''
        which = int(index) + (0 if kind == 'matrix' else 4)
        if coef is None:
            self.ctx.check(self.ctx.lib.b2_elemset_set_coefficient(self.elemset, which, None, 0))
            return
        coef = as_f64(coef).ravel()
        self.ctx.check(self.ctx.lib.b2_elemset_set_coefficient(self.elemset, which, coef.ctypes.data_as(c_vp), len(coef)))

    def set_coefficient_field(self, kind, index, field, power=1, scale=1.):
        '''solution-dependent coefficient scale * u_h^power of matrix form `index` (kind 'matrix') or vector form `index` (kind
        'vector'), u_h = sum_i field[i] N_i evaluated inside the kernel (b2_elemset_set_coefficient_field).  `field` is a device
        vector (DeviceBuffer / torch tensor / pointer) that the caller may update in place between assemblies; None removes it.'''
        which = int(index) + (0 if kind == 'matrix' else 4)
        self._keep = [k for k in getattr(self, '_keep', []) if k[0] != which]
        if field is None:
            self.ctx.check(self.ctx.lib.b2_elemset_set_coefficient_field(self.elemset, which, None, 0, 1.))
            return
        self._keep.append((which, field))
        self.ctx.check(self.ctx.lib.b2_elemset_set_coefficient_field(self.elemset, which, _devptr(field), int(power), float(scale)))

    def evaluate(self, fields=(), x=True, weights=True, values=True, grads=False):
        '''Sample.eval on the device (b2_evaluate_elemset_device): returns a dict with the requested arrays in point order:
        'x' [npoints, ndims], 'weights' [npoints] (w |det J|), 'values' [npoints, nfields, ncomp] and 'grads'
        [npoints, nfields, ncomp, ndims] of the discrete fields with coefficient vectors `fields` (each of length ndofs).'''
        fields = [as_f64(f).ravel() for f in fields]
        if any(len(f) != self.ndofs for f in fields):
            raise ValueError('coefficient vectors must have length ndofs')
        nf, nc, nd = len(fields), self.ncomp, self.ndims
        npts = self.npoints if self.quad is None else self.nsel * int(numpy.prod([len(r[0]) for r in self.rules]))
        ctx = self.ctx
        bufs, out = {}, {}
        coef = None
        if nf:
            coef = ctx.device_alloc(8 * nf * self.ndofs)
            coef.from_host(numpy.concatenate(fields))
        shapes = {'x': (npts, nd), 'weights': (npts,), 'values': (npts, nf, nc), 'grads': (npts, nf, nc, nd)}
        want = {'x': x, 'weights': weights, 'values': values and nf, 'grads': grads and nf}
        for key, shp in shapes.items():
            if want[key]:
                bufs[key] = ctx.device_alloc(8 * max(int(numpy.prod(shp)), 1))
        ptr = lambda key: bufs[key].ptr if key in bufs else None
        ctx.check(ctx.lib.b2_evaluate_elemset_device(ctx.handle, self.elemset, self.quad, self.geom, nf, None if coef is None else coef.ptr,
                                                     ptr('x'), ptr('weights'), ptr('values'), ptr('grads')))
        for key, buf in bufs.items():
            n = int(numpy.prod(shapes[key]))
            out[key] = buf.to_host()[:n].reshape(shapes[key]) if n else numpy.zeros(shapes[key])
        return out

    def _make_basis(self, bases, ncomp):
        h = _create_basis(self.ctx, bases, ncomp)
        self._fin.append(weakref.finalize(self, self.ctx.lib.b2_basis_destroy, h))
        return h

    csr_pattern = Plan.csr_pattern
    csr_pattern_device = Plan.csr_pattern_device
    spmv_device = Plan.spmv_device
    diagonal_device = Plan.diagonal_device
    cg_device = Plan.cg_device
    row_offset = Plan.row_offset
    _form_args = Plan._form_args

    def assemble_host(self, Ds=(), Cs=()):
        'zero-fill, integrate the selected elements, copy to host: returns ([values...], [rhs...]) (b2_assemble_elemset_host)'
        Ds, Cs, pD, pC = self._form_args(Ds, Cs)
        vals = [numpy.empty(self.nnz) for _ in Ds]
        rhs = [numpy.empty(self.ndofs) for _ in Cs]
        pv = (c_vp * max(len(vals), 1))(*[v.ctypes.data_as(c_vp) for v in vals])
        pr = (c_vp * max(len(rhs), 1))(*[r.ctypes.data_as(c_vp) for r in rhs])
        self.ctx.check(self.ctx.lib.b2_assemble_elemset_host(self.ctx.handle, self.pattern, self.elemset, self.quad, self.geom, len(Ds), pD, pv, len(Cs), pC, pr))
        return vals, rhs

    def assemble_device(self, Ds=(), Cs=(), values=(), rhs=(), sel_range=None):
        'accumulate the selected elements [sel_range) into device arrays (b2_assemble_elemset_device); asynchronous'
        Ds, Cs, pD, pC = self._form_args(Ds, Cs)
        assert len(values) == len(Ds) and len(rhs) == len(Cs)
        s0, s1 = sel_range if sel_range is not None else (0, -1)
        pv = (c_vp * max(len(values), 1))(*[_devptr(v) for v in values])
        pr = (c_vp * max(len(rhs), 1))(*[_devptr(r) for r in rhs])
        self.ctx.check(self.ctx.lib.b2_assemble_elemset_device(self.ctx.handle, self.pattern, self.elemset, self.quad, self.geom, c_i64(s0), c_i64(s1),
                                                               len(Ds), pD, pv, len(Cs), pC, pr))


class CsrPattern:
    '''An explicit CSR pattern on the device (b2_pattern_create_csr): what ``matrix.assemble_csr(values, rowptr, colidx, ncols)``
    receives (src/nutils/matrix/__init__.py:30-70).  Carries the device-resident matrix operations of a plan.'''

    def __init__(self, ctx, rowptr, colidx, ncols):
        self.ctx = ctx
        rowptr = numpy.ascontiguousarray(rowptr, dtype=numpy.int64)
        colidx = numpy.ascontiguousarray(colidx, dtype=numpy.int64)
        h = c_vp()
        ctx.check(ctx.lib.b2_pattern_create_csr(ctx.handle, len(rowptr) - 1, int(ncols), rowptr.ctypes.data_as(c_vp), colidx.ctypes.data_as(c_vp), ctypes.byref(h)))
        self.pattern = h
        self._fin = [weakref.finalize(self, ctx.lib.b2_pattern_destroy, h)]
        self.nnz = int(ctx.lib.b2_pattern_nnz(h))
        self.ndofs = int(ctx.lib.b2_pattern_nrows(h))
        self.ncols = int(ncols)
        self._csr = rowptr, colidx

    csr_pattern = Plan.csr_pattern
    csr_pattern_device = Plan.csr_pattern_device
    spmv_device = Plan.spmv_device
    diagonal_device = Plan.diagonal_device
    cg_device = Plan.cg_device
    row_offset = Plan.row_offset


class GeneralPlan:
    '''Assembly on a GENERAL dof map: simplex and mixed meshes, any basis given per element as tabulated
    reference data.  The CSR pattern is built on the device from the element dof lists (b2_pattern_general: key generation,
    radix sort, unique -- the reference's argsort / unique / compress_indices, evaluable.py:5646-5682), the element loop is
    b2_assemble_general_host.

    ndims, ncomp
    types   : list of dicts, one per element type: weights[nq], phi[nq, nfun], dphi[nq, nfun, ndims], gdphi[nq, nvert, ndims]
    etype   : int32[nelems] type of every element
    dofs    : list of int arrays (basis indices of the local functions of every element)  -- or (dofoff, dofs) arrays
    verts   : list of float arrays [nvert, ndims] (vertex coordinates of every element)   -- or (vertoff, coords) arrays
    nbasis  : number of basis functions
    '''

    def __init__(self, ctx, ndims, types, etype, dofs, verts, nbasis, ncomp=1):
        self.ctx = ctx
        self.ndims, self.ncomp = int(ndims), int(ncomp)
        self.etype = numpy.ascontiguousarray(etype, dtype=numpy.int32)
        self.nelems = len(self.etype)
        self.types = [dict(weights=as_f64(t['weights']), phi=as_f64(t['phi']), dphi=as_f64(t['dphi']), gdphi=as_f64(t['gdphi'])) for t in types]
        for t in self.types:
            nq, nf = t['phi'].shape
            if t['weights'].shape != (nq,) or t['dphi'].shape != (nq, nf, self.ndims) or t['gdphi'].ndim != 3 or t['gdphi'].shape[0] != nq or t['gdphi'].shape[2] != self.ndims:
                raise ValueError('inconsistent element tables')
        if isinstance(dofs, tuple):
            self.dofoff, self.dofs = (numpy.ascontiguousarray(a, dtype=numpy.int64) for a in dofs)
        else:
            self.dofoff = numpy.concatenate([[0], numpy.cumsum([len(d) for d in dofs])]).astype(numpy.int64)
            self.dofs = numpy.ascontiguousarray(numpy.concatenate([numpy.asarray(d, dtype=numpy.int64) for d in dofs]) if len(dofs) else numpy.zeros(0, dtype=numpy.int64))
        if isinstance(verts, tuple):
            self.vertoff, self.vertcoords = numpy.ascontiguousarray(verts[0], dtype=numpy.int64), as_f64(verts[1])
        else:
            self.vertoff = numpy.concatenate([[0], numpy.cumsum([len(v) for v in verts])]).astype(numpy.int64)
            self.vertcoords = as_f64(numpy.concatenate([numpy.asarray(v, dtype=float).reshape(-1, self.ndims) for v in verts]) if len(verts) else numpy.zeros((0, self.ndims)))
        for e in range(self.nelems):
            t = self.types[self.etype[e]]
            if self.dofoff[e + 1] - self.dofoff[e] != t['phi'].shape[1] or self.vertoff[e + 1] - self.vertoff[e] != t['gdphi'].shape[1]:
                raise ValueError('element {} does not match its type tables'.format(e))
        self.nbasis = int(nbasis)
        h = c_vp()
        ctx.check(ctx.lib.b2_pattern_general(ctx.handle, self.nelems, self.dofoff.ctypes.data_as(c_vp), self.dofs.ctypes.data_as(c_vp), self.dofoff.ctypes.data_as(c_vp),
                                             self.dofs.ctypes.data_as(c_vp), self.nbasis, self.nbasis, self.ncomp, ctypes.byref(h)))
        self.pattern = h
        self._fin = [weakref.finalize(self, ctx.lib.b2_pattern_destroy, h)]
        self.nnz = int(ctx.lib.b2_pattern_nnz(h))
        self.ndofs = int(ctx.lib.b2_pattern_nrows(h))
        self._csr = None

    csr_pattern = Plan.csr_pattern
    csr_pattern_device = Plan.csr_pattern_device
    spmv_device = Plan.spmv_device
    diagonal_device = Plan.diagonal_device
    cg_device = Plan.cg_device
    row_offset = Plan.row_offset
    _form_args = Plan._form_args

    def assemble_host(self, Ds=(), Cs=()):
        'zero-fill, integrate every element, copy to host: ([values...], [rhs...]) (b2_assemble_general_host)'
        Ds, Cs, pD, pC = self._form_args(Ds, Cs)
        vals = [numpy.empty(self.nnz) for _ in Ds]
        rhs = [numpy.empty(self.ndofs) for _ in Cs]
        pv = (c_vp * max(len(vals), 1))(*[v.ctypes.data_as(c_vp) for v in vals])
        pr = (c_vp * max(len(rhs), 1))(*[r.ctypes.data_as(c_vp) for r in rhs])
        i32 = lambda a: numpy.ascontiguousarray(a, dtype=numpy.int32)
        nq = i32([t['phi'].shape[0] for t in self.types])
        nf = i32([t['phi'].shape[1] for t in self.types])
        nv = i32([t['gdphi'].shape[1] for t in self.types])
        tab = lambda key: _lib.ptr_array([t[key] for t in self.types], _lib.p_f64)
        self.ctx.check(self.ctx.lib.b2_assemble_general_host(
            self.ctx.handle, self.pattern, self.ndims, self.ncomp, self.nelems, len(self.types), self.etype.ctypes.data_as(c_vp), nq.ctypes.data_as(c_vp),
            nf.ctypes.data_as(c_vp), nv.ctypes.data_as(c_vp), tab('weights'), tab('phi'), tab('dphi'), tab('gdphi'), tab('gdphi'), self.dofoff.ctypes.data_as(c_vp),
            self.dofs.ctypes.data_as(c_vp), self.vertoff.ctypes.data_as(c_vp), self.vertcoords.ctypes.data_as(c_vp), len(Ds), pD, pv, len(Cs), pC, pr))
        return vals, rhs


# ---- coefficient tensors of the north-star forms -------------------------------------------------------

def form_mass(ndims, ncomp=1):
    'D for int N_i N_j (per component)'
    D = numpy.zeros((ncomp, ndims + 1, ncomp, ndims + 1))
    for c in range(ncomp):
        D[c, 0, c, 0] = 1.
    return D


def form_stiffness(ndims, ncomp=1, conductivity=None):
    'D for int grad N_i . K grad N_j (K = identity by default)'
    K = numpy.eye(ndims) if conductivity is None else numpy.asarray(conductivity, dtype=float)
    D = numpy.zeros((ncomp, ndims + 1, ncomp, ndims + 1))
    for c in range(ncomp):
        D[c, 1:, c, 1:] = K
    return D


def form_elasticity(ndims, lmbda, mu, scale=1.):
    '''D for scale * int eps(v):sigma(u), sigma = lmbda tr(eps) I + 2 mu eps:
    D[c,1+k,e,1+l] = scale (lmbda d_ck d_el + mu (d_ce d_kl + d_cl d_ek)).'''
    eye = numpy.eye(ndims)
    D = numpy.zeros((ndims, ndims + 1, ndims, ndims + 1))
    D[:, 1:, :, 1:] = scale * (lmbda * numpy.einsum('ck,el->ckel', eye, eye) + mu * (numpy.einsum('ce,kl->ckel', eye, eye) + numpy.einsum('cl,ek->ckel', eye, eye)))
    return D


def form_load(ndims, ncomp=1, f=None):
    'C for int f_c N_i (f = ones by default)'
    C = numpy.zeros((ncomp, ndims + 1))
    C[:, 0] = 1. if f is None else numpy.asarray(f, dtype=float)
    return C
