'''Device timing of the element-set path on a synthetic finite-cell workload (development helper): a ball cut out of an
n^3 grid, octree quadrature of depth L on the cut elements (own generator -- NOT the reference's mosaic; parity of the
path is pinned by the goldens).  usage: python scripts/time_elemset.py [n] [degree] [L] [mma]'''
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from nutils_b200 import bspline, points, engine


from nutils_b200.fcm import octree_ball  # noqa: E402,F401


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
    ctx.set_option('elemset_queue', int(os.environ.get('ES_QUEUE', 1)))
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
