'''World-size-2 tests of the multi-GPU decompositions on CPU (gloo): the index arithmetic of nutils_b200.distributed and the
neighbour exchange, with the C oracle standing in for the per-rank integrator (the CUDA kernels are covered by the -m gpu tests:
test_rows_plane_ranges / test_element_ranges_accumulate exercise the same contracts on one GPU).'''

import os
import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util
from oracle import c_oracle, fem_oracle
from nutils_b200 import distributed, points


def _problem():
    nelems, degree = (5, 3, 4), 2
    b1 = util.bases_1d(nelems, degree, 'spline')
    rules = points.tensor_gauss(3, 2 * degree)
    rng = numpy.random.RandomState(3)
    X = numpy.stack(numpy.meshgrid(*[numpy.arange(n + 1.) for n in nelems], indexing='ij'))
    X = X + .2 * (rng.rand(*X.shape) - .5)
    prob = fem_oracle.Problem(nelems, [degree] * 3, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], X)
    return prob, b1


def _slot_lookup(rowptr, colidx, rows, cols):
    'CSR slot of every (row, col) pair'
    slots = numpy.empty(len(rows), dtype=numpy.int64)
    for k, (r, c) in enumerate(zip(rows, cols)):
        a, b = rowptr[r], rowptr[r + 1]
        slots[k] = a + numpy.searchsorted(colidx[a:b], c)
    return slots


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        prob, b1 = _problem()
        (K_full, rowptr, colidx), = c_oracle.assemble(prob, [('stiffness',)], [])[0]
        f_full = c_oracle.assemble(prob, [], [('load',)])[1][0]
        row_offset = lambda r: int(rowptr[r])
        per_layer = prob.ntotal // prob.nelems[0]
        if mode == 'planes':
            lay = distributed.PlaneLayout(b1, 1, rank, world, row_offset)
            e0, e1 = lay.elem_layers
            vals, rows, cols, rhs = c_oracle.element_loop(prob, [('stiffness',)], [('load',)], elem_range=(e0 * per_layer, e1 * per_layer))
            keep = (rows >= lay.row_lo) & (rows < lay.row_hi)   # the kernel's contract: only the owned rows are written
            win = numpy.zeros(lay.nvalues)
            numpy.add.at(win, _slot_lookup(rowptr, colidx, rows[keep], cols[keep]) - lay.off_lo, vals[0][keep])
            fwin = rhs[0][lay.row_lo:lay.row_hi]
            # no collective on the data path: the windows of the ranks tile the global arrays
            ok = numpy.allclose(win, K_full[lay.off_lo:lay.off_hi], rtol=1e-13, atol=1e-15)
            # the load vector of the overlap layers is incomplete outside the owned rows, complete inside
            ok &= numpy.allclose(fwin, f_full[lay.row_lo:lay.row_hi], rtol=1e-13, atol=1e-15)
            gathered = [None] * world
            dist.all_gather_object(gathered, (lay.off_lo, lay.off_hi, lay.row_lo, lay.row_hi))
            if rank == 0:
                ok &= gathered[0][0] == 0 and gathered[-1][1] == len(K_full) and all(a[1] == b[0] for a, b in zip(gathered[:-1], gathered[1:]))
                ok &= gathered[0][2] == 0 and gathered[-1][3] == prob.ndofs and all(a[3] == b[2] for a, b in zip(gathered[:-1], gathered[1:]))
        else:
            lay = distributed.SlabLayout(b1, 1, rank, world, row_offset)
            vals, rows, cols, rhs = c_oracle.element_loop(prob, [('stiffness',)], [('load',)], elem_range=lay.elem_range)
            assert rows.min() >= lay.row_lo and rows.max() < lay.row_hi
            win = numpy.zeros(lay.nvalues)
            numpy.add.at(win, _slot_lookup(rowptr, colidx, rows, cols) - lay.off_lo, vals[0])
            tw, tf = torch.from_numpy(win), torch.from_numpy(rhs[0][lay.row_lo:lay.row_hi].copy())
            distributed.exchange_interfaces(lay, [tw], [tf])
            ok = numpy.allclose(tw.numpy(), K_full[lay.off_lo:lay.off_hi], rtol=1e-13, atol=1e-15)
            ok &= numpy.allclose(tf.numpy(), f_full[lay.row_lo:lay.row_hi], rtol=1e-13, atol=1e-15)
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['planes', 'slabs'])
def test_world2_gloo(mode):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29400 + os.getpid() % 500 + (7 if mode == 'slabs' else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_plane_layout_arithmetic():
    prob, b1 = _problem()
    (K, rowptr, colidx), = c_oracle.assemble(prob, [('mass',)], [])[0]
    for world in (1, 2, 3, 7):
        lays = [distributed.PlaneLayout(b1, 1, r, world, lambda row: int(rowptr[row])) for r in range(world)]
        assert lays[0].row_lo == 0 and lays[-1].row_hi == prob.ndofs
        assert all(a.row_hi == b.row_lo and a.off_hi == b.off_lo for a, b in zip(lays[:-1], lays[1:]))
        assert sum(l.nvalues for l in lays) == len(K)
        for l in lays:
            p0, p1 = l.plane_range
            assert l.elem_layers == (max(0, p0 - 2), min(prob.nelems[0], p1))
    with pytest.raises(ValueError):
        distributed.PlaneLayout(b1, 1, 0, 100, lambda row: 0)


# ---- element sets (trimmed topologies): cost-balanced element ranges + the same neighbour exchange ---------------

def _cut_problem():
    from tests.test_gpu_elemset import _random_cut
    return _random_cut(11, (8, 3, 3), 2, keep=.7, maxpts=30)


def _elemset_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        prob = _cut_problem()
        b1 = util.bases_1d(prob.nelems, prob.degree[0], 'spline')
        mats, vecs = fem_oracle.assemble(prob, [('stiffness',)], [('load',)])
        K_full, rowptr, colidx = mats[0]
        lay = distributed.ElemSetLayout(b1, 1, rank, world, lambda r: int(rowptr[r]), elem_ids=prob.elem_ids, qoff=prob.qoff,
                                        renumber=prob.renumber, nbasis_new=prob.nbasis_new)
        win = numpy.zeros(lay.nvalues)
        fwin = numpy.zeros(lay.nrows)
        for isel in range(*lay.sel_range):     # the per-rank integrator: what b2_assemble_elemset_device(sel_begin, sel_end) does
            dofs, N, grad, wdet = fem_oracle.element_data(prob, isel)
            blk = fem_oracle.element_matrix(('stiffness',), N, grad, wdet)
            assert dofs.min() >= lay.row_lo and dofs.max() < lay.row_hi
            slots = _slot_lookup(rowptr, colidx, numpy.repeat(dofs, len(dofs)), numpy.tile(dofs, len(dofs)))
            numpy.add.at(win, slots - lay.off_lo, blk.ravel())
            numpy.add.at(fwin, dofs - lay.row_lo, fem_oracle.element_vector(('load',), N, grad, wdet))
        tw, tf = torch.from_numpy(win), torch.from_numpy(fwin)
        distributed.exchange_interfaces(lay, [tw], [tf])
        ok = numpy.allclose(tw.numpy(), K_full[lay.off_lo:lay.off_hi], rtol=1e-12, atol=1e-15)
        ok &= numpy.allclose(tf.numpy(), vecs[0][lay.row_lo:lay.row_hi], rtol=1e-12, atol=1e-15)
        ok &= abs(lay.cost_share - 1 / world) < .2
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put(int(flag))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_elemset():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29400 + os.getpid() % 500 + 13
    procs = [ctx.Process(target=_elemset_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1


def test_balanced_cuts():
    cost = numpy.array([1., 1, 1, 1, 100, 1, 1, 1])
    cuts = distributed.balanced_cuts(cost, 2)
    assert cuts[0] == 0 and cuts[-1] == 8 and 4 <= cuts[1] <= 5
    assert list(distributed.balanced_cuts(numpy.ones(9), 3)) == [0, 3, 6, 9]
