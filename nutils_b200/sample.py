'''Sample: quadrature points on a topology; entry point of the accelerated path.

Mirrors ``nutils.sample.Sample.integrate`` / ``integral`` (src/nutils/sample.py:160-190) and
``_Integral`` (sample.py:944-956).  Evaluation builds (once per sample/basis/geometry) a
device-resident :class:`engine.Plan` and launches the CUDA assembly through the C ABI.
'''

import numpy
from . import function, engine


class Sample:
    def __init__(self, topo, rules, device=0):
        self.topo = topo
        self.rules = rules
        self.ndims = topo.ndims
        self.nelems = len(topo)
        self.npoints = self.nelems * int(numpy.prod([len(r[0]) for r in rules]))
        self.device = device
        self._plans = {}

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

    def integrate_device(self, funcs, arguments=None):
        '''like integrate_sparse, but the values of 2-D integrands STAY in HBM: matrices come back as matrix.DeviceMatrix
        (products and constrained CG solves on the device), vectors as numpy arrays.'''
        from . import matrix
        if arguments:
            raise NotImplementedError('integrands with arguments are outside the accelerated path')
        single = not isinstance(funcs, (tuple, list))
        integrals = [self.integral(f) for f in ((funcs,) if single else funcs)]
        res = [None] * len(integrals)
        groups = {}
        for k, integral in enumerate(integrals):
            groups.setdefault((id(integral.func.space), id(integral.func.jac)), []).append(k)
        for ks in groups.values():
            plan = self.plan(integrals[ks[0]].func.space, integrals[ks[0]].func.jac)
            mats = [k for k in ks if integrals[k].kind == 'matrix']
            vecs = [k for k in ks if integrals[k].kind == 'vector']
            vbufs = [plan.ctx.device_alloc(8 * plan.nnz) for _ in mats]
            rbufs = [plan.ctx.device_alloc(8 * plan.ndofs) for _ in vecs]
            plan.assemble_rows_device([integrals[k].tensor for k in mats], [integrals[k].tensor for k in vecs], vbufs, rbufs)
            for k, buf in zip(mats, vbufs):
                res[k] = matrix.DeviceMatrix(plan, buf)
            for k, buf in zip(vecs, rbufs):
                res[k] = buf.to_host()
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
            entry = self._plans[key] = plan, geom.nodes
        return entry[0]

    def _evaluate(self, integrals):
        '''Assemble a list of Integrals that share this sample and one basis with ONE launch per geometry.

        Returns per integral: the rhs vector, or (values, rowptr, colidx).'''
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
