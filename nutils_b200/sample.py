'''Sample: quadrature points on a topology; entry point of the accelerated path.

Mirrors ``nutils.sample.Sample.integrate`` / ``integral`` (src/nutils/sample.py:160-190) and
``_Integral`` (sample.py:944-956).  Evaluation builds (once per sample/basis/geometry) a
device-resident :class:`engine.Plan` and launches the CUDA assembly through the C ABI.
'''

import numpy
from . import function, engine


class Sample:
    def __init__(self, topo, rules, device=0, faces=None):
        self.topo = topo
        self.rules = rules
        self.ndims = topo.ndims
        self.faces = faces  # None: volume sample; tuple of (direction, side): boundary sample on those sides
        self.nelems = len(topo)
        self.npoints = self.nelems * int(numpy.prod([len(r[0]) for r in rules]))
        self.device = device
        self._plans = {}
        self._esplans = {}

    # -- reference API -------------------------------------------------------------------------------

    def integral(self, func):
        'postponed integration: ``sample.integrate(f) == sample.integral(f).eval()`` (sample.py:177-190)'
        return function.Integral(self, function.Array.cast(func))

    def integrate(self, funcs, arguments=None):
        'integrate one function or a tuple of functions (sample.py:160-175)'
        if isinstance(funcs, (tuple, list)):
            return function.eval(tuple(self.integral(f) for f in funcs), arguments or None)
        return function.eval(self.integral(funcs), arguments or None)

    def integrate_sparse(self, funcs, arguments=None):
        'like integrate, but 2-D integrands come back as matrix.Matrix objects (CSR) instead of dense arrays'
        from . import matrix
        single = not isinstance(funcs, (tuple, list))
        integrals = [self.integral(f) for f in ((funcs,) if single else funcs)]
        outs = function.eval(tuple(function.as_csr(i) if i.kind == 'matrix' else i for i in integrals), arguments or None)
        res = tuple(matrix.assemble_csr(*o, ncols=len(o[1]) - 1) if i.kind == 'matrix' else o for i, o in zip(integrals, outs))
        return res[0] if single else res

    def integrate_device(self, funcs, arguments=None, out=None):
        '''like integrate_sparse, but the values of 2-D integrands STAY in HBM: matrices come back as matrix.DeviceMatrix
        (products and constrained CG solves on the device), vectors as numpy arrays.

        out : the result of a previous call with the same integrands; its device buffers are written again instead of
              allocating new ones (a re-assembly inside a time or Newton loop), vectors then stay on the device as
              engine.DeviceBuffer objects (``.to_host()`` copies them).'''
        from . import matrix
        if arguments:
            raise NotImplementedError('integrands with arguments are outside the accelerated path')
        single = not isinstance(funcs, (tuple, list))
        integrals = [self.integral(f) for f in ((funcs,) if single else funcs)]
        if self.faces is not None or any(i.func.coef is not None for i in integrals):
            # boundary samples and pointwise coefficients run on the element-set route, which assembles through host buffers
            raise NotImplementedError('integrate_device covers volume integrals with constant coefficients; use integrate_sparse for boundary '
                                      'samples and integrands with coefficient functions')
        prev = None if out is None else ([out] if single else list(out))
        if prev is not None and len(prev) != len(integrals):
            raise ValueError('out does not match the integrands')
        res = [None] * len(integrals)
        groups = {}
        for k, integral in enumerate(integrals):
            groups.setdefault((id(integral.func.space), id(integral.func.jac)), []).append(k)
        for ks in groups.values():
            plan = self.plan(integrals[ks[0]].func.space, integrals[ks[0]].func.jac)
            mats = [k for k in ks if integrals[k].kind == 'matrix']
            vecs = [k for k in ks if integrals[k].kind == 'vector']
            if prev is None:
                vbufs = [plan.ctx.device_alloc(8 * plan.nnz) for _ in mats]
                rbufs = [plan.ctx.device_alloc(8 * plan.ndofs) for _ in vecs]
            else:
                vbufs = [prev[k].values for k in mats]
                rbufs = [prev[k] for k in vecs]
                if any(getattr(prev[k], 'plan', None) is not plan for k in mats) or any(not isinstance(b, engine.DeviceBuffer) or b.nbytes != 8 * plan.ndofs for b in rbufs):
                    raise ValueError('out does not match the integrands')
            plan.assemble_rows_device([integrals[k].tensor for k in mats], [integrals[k].tensor for k in vecs], vbufs, rbufs)
            for k, buf in zip(mats, vbufs):
                res[k] = prev[k] if prev is not None else matrix.DeviceMatrix(plan, buf)
            for k, buf in zip(vecs, rbufs):
                res[k] = buf if prev is not None else buf.to_host()
        return res[0] if single else tuple(res)

    def integrate_device_buffers(self, funcs):
        'first call of a re-assembly loop: like integrate_device, but vectors stay on the device too (engine.DeviceBuffer); pass the result as ``out=``'
        single = not isinstance(funcs, (tuple, list))
        integrals = [self.integral(f) for f in ((funcs,) if single else funcs)]
        res = self.integrate_device(funcs)
        res = [res] if single else list(res)
        for k, integral in enumerate(integrals):
            if integral.kind == 'vector':
                plan = self.plan(integral.func.space, integral.func.jac)
                buf = plan.ctx.device_alloc(8 * plan.ndofs)
                buf.from_host(res[k])
                res[k] = buf
        return res[0] if single else tuple(res)

    # -- engine --------------------------------------------------------------------------------------

    def plan(self, space, geom):
        key = id(space), id(geom)
        entry = self._plans.get(key)
        if entry is None or entry[1] is not geom.nodes:
            ctx = engine.Context.get(self.device)
            if entry is not None and entry[0].nodes.shape == geom.nodes.shape:
                entry[0].update_nodes(geom.nodes)
                plan = entry[0]
            else:
                plan = engine.Plan(ctx, space.bases1d, self.rules, geom.nodes, ncomp=space.ncomp)
            entry = self._plans[key] = plan, geom.nodes, space, geom   # space and geom are held so that their ids cannot be reused
        return entry[0]

    # -- element-set route: boundary samples and integrands with coefficient functions ---------------

    def _face_tables(self, dim, side):
        'adjacent volume elements of one side and the face Gauss points in their local coordinates'
        shape = self.topo.shape
        nd = self.ndims
        grids = [numpy.arange(n) for n in shape]
        grids[dim] = numpy.array([shape[dim] - 1 if side else 0])
        idx = numpy.stack(numpy.meshgrid(*grids, indexing='ij'), -1).reshape(-1, nd)
        elem_ids = numpy.sort(numpy.ravel_multi_index(idx.T, shape))
        tang = [d for d in range(nd) if d != dim]
        xi = numpy.zeros((1, nd))
        w = numpy.ones(1)
        xi[:, dim] = float(side)
        for d in tang:  # tensor product of the 1-D rules of the tangential directions
            x1, w1 = self.rules[d]
            xi = numpy.repeat(xi, len(x1), axis=0)
            xi[:, d] = numpy.tile(x1, len(w))
            w = numpy.repeat(w, len(x1)) * numpy.tile(w1, len(w))
        return elem_ids, xi, w

    def _physical_points(self, nodes, elem_ids, xi):
        'x[nsel, nq, ndims] of the multilinear geometry at local points xi[nq, ndims] of the elements elem_ids'
        nd = self.ndims
        idx = numpy.unravel_index(elem_ids, self.topo.shape)
        x = numpy.zeros((len(elem_ids), len(xi), nd))
        for corner in range(1 << nd):
            bits = [(corner >> (nd - 1 - d)) & 1 for d in range(nd)]
            phi = numpy.ones(len(xi))
            for d, b in enumerate(bits):
                phi = phi * (xi[:, d] if b else 1. - xi[:, d])
            X = nodes[(slice(None),) + tuple(i + b for i, b in zip(idx, bits))]  # (ndims, nsel)
            x += phi[None, :, None] * X.T[:, None, :]
        return x

    def _esplan(self, space, geom, face):
        'engine.ElemSetPlan of this sample for one side (or the volume: face None), cached; returns (plan, elem_ids, xi)'
        key = id(space), id(geom), face
        entry = self._esplans.get(key)
        if entry is None or entry[1] is not geom.nodes:
            ctx = engine.Context.get(self.device)
            if face is None:
                xi = numpy.stack([g.ravel() for g in numpy.meshgrid(*[r[0] for r in self.rules], indexing='ij')], -1)
                elem_ids = numpy.arange(self.nelems)
                plan = engine.ElemSetPlan(ctx, space.bases1d, nodes=geom.nodes, ncomp=space.ncomp, rules=self.rules)
            else:
                elem_ids, xi, w = self._face_tables(*face)
                nq = len(w)
                qoff = numpy.arange(len(elem_ids) + 1, dtype=numpy.int64) * nq
                plan = engine.ElemSetPlan(ctx, space.bases1d, nodes=geom.nodes, ncomp=space.ncomp, elem_ids=elem_ids, qoff=qoff,
                                          qcoords=numpy.tile(xi, (len(elem_ids), 1)), qweights=numpy.tile(w, len(elem_ids)))
                plan.set_faces(numpy.full(len(elem_ids), face[0], dtype=numpy.int8))
            entry = self._esplans[key] = plan, geom.nodes, elem_ids, xi, space, geom
        return entry[0], entry[2], entry[3]

    def eval_fields(self, space, geom, fields=(), grads=False):
        '''``Sample.eval`` for discrete fields (sample.py:192-215): a dict with the physical coordinates 'x' [npoints, ndims],
        the integration weights 'weights' [npoints] (w |det J|; on a boundary sample the surface measure), the 'values'
        [npoints, nfields, ncomp] of u_f = sum_i fields[f][i] N_i and, on request, their physical 'grads'.  Points are
        ordered like the reference orders them (elements, then the tensor points in C order).  With these an integral of
        any pointwise function of the solution is ``(weights * g(x, values)).sum()``.'''
        faces = self.faces if self.faces is not None else (None,)
        outs = []
        for face in faces:
            plan, _, _ = self._esplan(space, geom, face)
            outs.append(plan.evaluate(list(fields), grads=grads))
        return {k: numpy.concatenate([o[k] for o in outs]) for k in outs[0]}

    def _evaluate_elemset(self, integrals):
        'boundary integrals and integrands with coefficient functions: engine.ElemSetPlan, one plan per side'
        from . import matrix as _matrix
        geom = integrals[0].func.jac
        space = integrals[0].func.space
        if any(i.func.jac is not geom for i in integrals):
            raise NotImplementedError('integrals over different geometries in one evaluation')
        faces = self.faces if self.faces is not None else (None,)
        mats = [k for k, i in enumerate(integrals) if i.kind == 'matrix']
        vecs = [k for k, i in enumerate(integrals) if i.kind == 'vector']
        if len(faces) > 1 and mats:
            raise NotImplementedError('matrix-valued integrals over several sides at once (integrate the sides separately)')
        if len(mats) > 4 or len(vecs) > 4:
            raise NotImplementedError('more than four matrices or vectors per evaluation')
        out = [None] * len(integrals)
        ctx = engine.Context.get(self.device)
        for face in faces:
            plan, elem_ids, xi = self._esplan(space, geom, face)
            xq = None
            for slot, ks in (('matrix', mats), ('vector', vecs)):
                for j, k in enumerate(ks):
                    coef = integrals[k].func.coef
                    if coef is None:
                        plan.set_coefficient(slot, j, None)
                    else:
                        if xq is None:
                            xq = self._physical_points(geom.nodes, elem_ids, xi)
                        plan.set_coefficient(slot, j, numpy.ascontiguousarray(coef(xq), dtype=float))
            values, rhs = plan.assemble_host([integrals[k].tensor for k in mats], [integrals[k].tensor for k in vecs])
            if mats:
                rowptr, colidx = plan.csr_pattern()
            for k, v in zip(mats, values):
                out[k] = v, rowptr, colidx
            for k, r in zip(vecs, rhs):
                out[k] = r if out[k] is None else out[k] + r
        return out

    def _evaluate(self, integrals):
        '''Assemble a list of Integrals that share this sample and one basis with ONE launch per geometry.

        Returns per integral: the rhs vector, or (values, rowptr, colidx).'''
        if self.faces is not None or any(i.func.coef is not None for i in integrals):
            return self._evaluate_elemset(integrals)
        out = [None] * len(integrals)
        bygeom = {}
        for k, integral in enumerate(integrals):
            bygeom.setdefault(id(integral.func.jac), []).append(k)
        for ks in bygeom.values():
            geom = integrals[ks[0]].func.jac
            space = integrals[ks[0]].func.space
            plan = self.plan(space, geom)
            mats = [k for k in ks if integrals[k].kind == 'matrix']
            vecs = [k for k in ks if integrals[k].kind == 'vector']
            if len(mats) > 4 or len(vecs) > 4:
                raise NotImplementedError('more than four matrices or vectors per evaluation')
            values, rhs = plan.assemble_host([integrals[k].tensor for k in mats], [integrals[k].tensor for k in vecs])
            if mats:
                rowptr, colidx = plan.csr_pattern()
            for k, v in zip(mats, values):
                out[k] = v, rowptr, colidx
            for k, r in zip(vecs, rhs):
                out[k] = r
        return out
