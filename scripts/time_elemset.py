'''Device timing of the element-set path on a synthetic finite-cell workload (development helper): a ball cut out of an
n^3 grid, octree quadrature of depth L on the cut elements (own generator -- NOT the reference's mosaic; parity of the
path is pinned by the goldens).  usage: python scripts/time_elemset.py [n] [degree] [L] [mma]'''
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from nutils_b200 import bspline, points, engine


def octree_ball(n, degree, L, radius=.8):
    'elem_ids, qoff, qcoords, qweights, renumber, nbasis_new for the ball |x| < radius in [-1,1]^3'
    v = numpy.linspace(-1, 1, n + 1)
    gx, gw = points.gauss1(2 * degree)
    g3 = numpy.stack(numpy.meshgrid(gx, gx, gx, indexing='ij'), -1).reshape(-1, 3)
    w3 = numpy.einsum('i,j,k->ijk', gw, gw, gw).ravel()
    X = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'), -1)
    inside = numpy.linalg.norm(X, axis=-1) < radius
    c = [inside[i:n + i, j:n + j, k:n + k] for i in (0, 1) for j in (0, 1) for k in (0, 1)]
    nin = numpy.sum(c, axis=0)
    full = nin == 8
    cut = (nin > 0) & ~full
    # a cell without inside vertices may still touch the ball: ignored (thin slivers), as a midpoint rule would
    m = 2 ** L
    sub = (numpy.stack(numpy.meshgrid(*[numpy.arange(m)] * 3, indexing='ij'), -1).reshape(-1, 1, 3) + g3[None]) / m   # [m^3, nq, 3]
    subc = (numpy.stack(numpy.meshgrid(*[numpy.arange(m)] * 3, indexing='ij'), -1).reshape(-1, 3) + .5) / m          # sub-cell centres
    elem_ids, coords, weights, qoff = [], [], [], [0]
    h = 2. / n
    for e in numpy.flatnonzero((full | cut).ravel()):
        i, j, k = numpy.unravel_index(e, (n, n, n))
        if full[i, j, k]:
            coords.append(g3)
            weights.append(w3)
        else:
            org = numpy.array([v[i], v[j], v[k]])
            keep = numpy.linalg.norm(org + subc * h, axis=-1) < radius
            coords.append(sub[keep].reshape(-1, 3))
            weights.append(numpy.tile(w3 / m ** 3, keep.sum()))
        elem_ids.append(e)
        qoff.append(qoff[-1] + len(weights[-1]))
    elem_ids = numpy.array(elem_ids, dtype=numpy.int64)
    nd = n + degree
    used = numpy.zeros((nd, nd, nd), dtype=bool)
    ii, jj, kk = numpy.unravel_index(elem_ids, (n, n, n))
    for a in range(degree + 1):
        for b in range(degree + 1):
            for c_ in range(degree + 1):
                used[ii + a, jj + b, kk + c_] = True
    dofs = numpy.flatnonzero(used.ravel())
    renumber = numpy.full(nd ** 3, len(dofs), dtype=numpy.int64)
    renumber[dofs] = numpy.arange(len(dofs))
    return elem_ids, numpy.array(qoff, dtype=numpy.int64), numpy.concatenate(coords), numpy.concatenate(weights), renumber, len(dofs)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    mma = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if int(os.environ.get('WORLD_SIZE', 1)) > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ctx = engine.Context.get(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option('elemset_mma', mma)
    t0 = time.perf_counter()
    elem_ids, qoff, qc, qw, ren, nbn = octree_ball(n, p, L)
    t1 = time.perf_counter()
    b1 = [bspline.spline_basis_1d(n, p) for _ in range(3)]
    v = numpy.linspace(-1, 1, n + 1)
    nodes = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'))
    plan = engine.ElemSetPlan(ctx, b1, nodes=nodes, elem_ids=elem_ids, qoff=qoff, qcoords=qc, qweights=qw, renumber=ren, nbasis_new=nbn)
    ctx.synchronize()
    t2 = time.perf_counter()
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    dev = torch.device('cuda', local)
    Ds, Cs = [engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)]
    if world > 1:
        # torchrun: cost-balanced element ranges per rank, windows of the global CSR, ONE neighbour exchange (NCCL) after the kernel
        import torch.distributed as dist
        from nutils_b200 import distributed
        lay = distributed.ElemSetLayout(b1, 1, rank, world, plan.row_offset, elem_ids=elem_ids, qoff=qoff, renumber=ren, nbasis_new=nbn)
        K = torch.zeros(lay.nvalues, dtype=torch.float64, device=dev)
        M = torch.zeros(lay.nvalues, dtype=torch.float64, device=dev)
        f = torch.zeros(lay.nrows, dtype=torch.float64, device=dev)
        ptrs = [K.data_ptr() - 8 * lay.off_lo, M.data_ptr() - 8 * lay.off_lo], [f.data_ptr() - 8 * lay.row_lo]

        def step():
            K.zero_(); M.zero_(); f.zero_()
            plan.assemble_device(Ds, Cs, ptrs[0], ptrs[1], sel_range=lay.sel_range)
            distributed.exchange_interfaces(lay, [K, M], [f])
    else:
        K = torch.zeros(plan.nnz, dtype=torch.float64, device=dev)
        M = torch.zeros(plan.nnz, dtype=torch.float64, device=dev)
        f = torch.zeros(plan.ndofs, dtype=torch.float64, device=dev)

        def step():
            K.zero_(); M.zero_(); f.zero_()
            plan.assemble_device(Ds, Cs, [K, M], [f])
    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        own = lay.own_rows
        sums = torch.stack([M[:plan.row_offset(lay.row_lo + own.stop) - lay.off_lo].sum(), f[own].sum()])  # owned rows only: shared rows are complete on both sides
        dist.all_reduce(sums)
        vol, fsum = float(sums[0]), float(sums[1])
    else:
        vol, fsum = float(M.sum()), float(f.sum())
    if rank == 0:
        print(json.dumps({'workload': 'finite-cell ball in {}^3, p={}, octree depth {}'.format(n, p, L), 'n_gpus': world, 'mma': mma, 'kept_elements': len(elem_ids), 'points': int(qoff[-1]),
                          'max_points_per_element': int(numpy.diff(qoff).max()), 'ndofs': plan.ndofs, 'nnz': plan.nnz, 'ms': ms, 'dof_per_s': plan.ndofs / ms * 1e3,
                          'points_per_s': int(qoff[-1]) / ms * 1e3, 'volume': vol, 'exact_volume': 4 / 3 * numpy.pi * .8 ** 3, 'sumM_minus_sumf': vol - fsum,
                          'host_quadrature_s': t1 - t0, 'plan_s (upload + pattern)': t2 - t1}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
