'''Linear systems on the accelerated path: the part of ``nutils.solver.System`` (src/nutils/solver.py:199-263, 318-425,
562-612) that a linear problem with constant or pointwise coefficients needs, with the matrix resident in HBM.

The reference builds residual and jacobian from ONE functional by symbolic differentiation; that machinery (expression
parser, autodiff, evaluable graph) is outside the scope of this repository.  Here the bilinear and the linear form are
given explicitly, which for the linear examples is the same information:

    K = topo.integral(<grad v . grad u> * J, degree)          # 2-D integrand  -> jacobian
    f = topo.boundary['right'].integral(<v g> * J, degree)    # 1-D integrands -> minus the residual at u = 0
    cons = solve_constraints([(B_left, None), (B_top, b_top)])
    u = LinearSystem(K, [f]).solve(constrain=cons)
'''

import numpy
from . import function


def solve_constraints(terms, droptol=1e-12):
    '''Minimise sum_k int_Gk (u - u_k)^2 dS over the dofs that contribute: the boundary projection of
    ``System.solve_constraints`` (solver.py:562-612).

    terms : sequence of (mass_integral, load_integral or None): int N_i N_j dS and int N_i u_k dS over one boundary each.
    Returns the constraint vector: a number for every dof whose diagonal entry of the summed mass matrix exceeds `droptol`,
    NaN for the free dofs (the convention of ``Matrix.solve(constrain=...)``, matrix/_base.py:115-120).'''
    import scipy.sparse
    import scipy.sparse.linalg
    B = rhs = None
    for mass, load in terms:
        evaluated = function.eval([function.as_csr(mass)] + ([load] if load is not None else []))
        v, rp, ci = evaluated[0]
        n = len(rp) - 1
        Bk = scipy.sparse.csr_matrix((v, ci, rp), shape=(n, n))
        B = Bk if B is None else B + Bk
        bk = evaluated[1] if load is not None else numpy.zeros(n)
        rhs = bk if rhs is None else rhs + bk
    diag = B.diagonal()
    rows = numpy.flatnonzero(abs(diag) > droptol)
    cons = numpy.full(B.shape[0], numpy.nan)
    if len(rows):
        cons[rows] = scipy.sparse.linalg.spsolve(B[rows][:, rows].tocsc(), rhs[rows])
    return cons


class LinearSystem:
    '''jacobian integral + load integrals of a linear problem; assembled once on the device (the reference caches constant
    matrices the same way, ``is_constant_matrix``, solver.py:255-256, 323-325).'''

    def __init__(self, jacobian, loads=()):
        if not isinstance(jacobian, function.Integral) or jacobian.kind != 'matrix':
            raise NotImplementedError('the jacobian must be a 2-D integral on the accelerated path')
        self.jacobian = jacobian
        self.loads = list(loads)
        self._matrix = self._rhs = None

    def assemble(self):
        '(matrix.DeviceMatrix, rhs) -- ``System.assemble_jacobian_residual`` at u = 0 with rhs = -residual'
        if self._matrix is None:
            if self.jacobian.sample.faces is not None or self.jacobian.func.coef is not None:
                raise NotImplementedError('device-resident jacobians with boundary terms or coefficient functions; use function.eval(function.as_csr(...))')
            self._matrix = self.jacobian.sample.integrate_device(self.jacobian.func)
            n = self._matrix.shape[0]
            rhs = numpy.zeros(n)
            for load in self.loads:
                rhs = rhs + function.eval(load)
            self._rhs = rhs
        return self._matrix, self._rhs

    def solve(self, constrain=None, lhs0=None, atol=0., rtol=1e-12, maxiter=0):
        '``System.solve`` for a linear problem: Jacobi-PCG on the device with the reference constraint semantics'
        K, rhs = self.assemble()
        return K.solve(rhs, lhs0=lhs0, constrain=constrain, atol=atol, rtol=rtol, maxiter=maxiter)
