#!/usr/bin/env python
'''Benchmark of the element-integration hot path (BASELINE.json: assembled DOFs/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 128] [--degree 2]

Workload (BASELINE.json configs[1]): 3-D Poisson on a structured n^3 grid (default 128^3), degree-2
B-splines, Gauss degree 4 (27 points), stiffness K + mass M + load vector f in ONE pass, on the
GENERAL-geometry code path: a multilinear nodal geometry whose nodes are perturbed, so every
element has its own per-point Jacobians (no uniform-mesh shortcut).  A "step" = one owner-computes
assembly pass (b2_assemble_rows_device): every stored value of K, M and f is written exactly once,
so there is no zero-fill.  N>1 (torchrun): weak scaling on an (N n) x n x n mesh, rank r owns 1/N of
the dof planes along x and the CSR rows that go with them and integrates the two element layers
below its first plane again instead of communicating -- no collective on the data path
(nutils_b200.distributed.PlaneLayout).  --path scatter selects the element-scatter kernels
(zero-fill + atomics + neighbour exchange over NCCL) for comparison.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, device-timed (CUDA events, max over
ranks).  `e2e`: the same pass through the host-buffer C-ABI call (b2_assemble_host) including the
H2D copy of the nodal coordinates and the D2H copy of K, M, f into pinned host memory.
`roofline`: algorithmic bytes (8 B per stored CSR value + 8 B per rhs entry + 8 B per nodal
coordinate) over the assembly kernel's own average device time against the measured HBM peak.
`cpu_baseline`: the C restatement of the reference algorithm (oracle/, kind "port") on all host
cores on a bounded sample of the same workload.
'''

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'assembled DOFs/sec (stiffness+mass+RHS), 3D Poisson p=2'
UNIT = 'DOF/s'
METRIC_ELAST = 'assembled DOFs/sec (stiffness+RHS), 3D elasticity p=2'


def make_nodes(shape, seed=0, warp=.2):
    'perturbed nodal coordinates of the unit-spaced grid, float64[3, ...]; deterministic'
    rng = numpy.random.RandomState(seed)
    verts = [numpy.linspace(0, n / shape[-1], n + 1) for n in shape]
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    X += warp / shape[-1] * (rng.rand(*X.shape) - .5)
    return X


def workload_name(n, degree, N, elast=False):
    mesh = '{}x{}x{}'.format(n * N, n, n)
    if elast:
        return '3D linear elasticity {} p={} spline (3 components), gauss{} ({} pts), K+f, nodal multilinear geometry (per-point Jacobians)'.format(mesh, degree, 2 * degree, (degree + 1) ** 3)
    return '3D Poisson {} p={} spline, gauss{} ({} pts), K+M+f, nodal multilinear geometry (per-point Jacobians)'.format(mesh, degree, 2 * degree, (degree + 1) ** 3)


# ---- CPU arm: the oracle port on host cores ----------------------------------------------------------

def cpu_assembly_rate(n, degree, nthreads=0, repeats=1, elast=False):
    '''DOF/s of the C restatement of the reference algorithm (element loop on all cores + serial
    sort/unique/accumulate), on an n^3 sample of the workload.'''
    from nutils_b200 import bspline, points
    from oracle import fem_oracle, c_oracle
    b1 = [bspline.spline_basis_1d(n, degree) for _ in range(3)]
    rules = points.tensor_gauss(3, 2 * degree)
    prob = fem_oracle.Problem((n,) * 3, [degree] * 3, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], make_nodes((n,) * 3), ncomp=3 if elast else 1)
    mforms = [('elasticity', 1., .5 / .3 - 1.)] if elast else [('stiffness',), ('mass',)]
    vforms = [('generic', numpy.array([[0., 0, 0, 0], [0, 0, 0, 0], [-1., 0, 0, 0]]))] if elast else [('load',)]
    c_oracle.lib()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        c_oracle.assemble(prob, mforms, vforms, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return prob.ndofs / best, best, prob.ndofs, (nthreads or c_oracle.max_threads())


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    # bounded sample: ~1.2 s per step at 32^3 on 16 cores, ~0.5 s at 24^3 -- the whole run stays within a few minutes
    n = min(args.cpu_n, 32 if args.warmup + args.steps <= 40 else 24)
    times = []
    for i in range(args.warmup + args.steps):
        rate, dt, ndofs, cores = cpu_assembly_rate(n if args.workload == 'poisson' else min(n, 20), args.degree, elast=args.workload == 'elasticity')
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = ndofs / t
    sample = '{}^3 elements ({} dofs) of the workload, general geometry; C port of the reference algorithm: threaded element loop + serial stable sort/unique/accumulate'.format(n, ndofs)
    line = {
        'impl': 'reference', 'metric': (METRIC_ELAST if args.workload == 'elasticity' else METRIC).replace('p=2', 'p={}'.format(args.degree)), 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args.n, args.degree, args.gpus), 'timed_sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---- clocks --------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(.06)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smmax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smmax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None, 'sm_max_mhz': smmax, 'reasons': sorted(reasons), 'samples': len(sm)}


# ---- GPU arm -------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    from nutils_b200 import bspline, points, engine, distributed

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torchrun --nproc-per-node {} for --gpus {}'.format(args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n, p, N = args.n, args.degree, world
    shape = (n * N, n, n)
    ctx = engine.Context.get(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    b1 = [bspline.spline_basis_1d(m, p) for m in shape]
    rules = points.tensor_gauss(3, 2 * p)
    nodes = make_nodes(shape)
    elast = args.workload == 'elasticity'
    ncomp = 3 if elast else 1
    plan = engine.Plan(ctx, b1, rules, nodes, ncomp=ncomp)
    rows_path = args.path == 'rows'
    layout = (distributed.PlaneLayout if rows_path else distributed.SlabLayout)(b1, ncomp, rank, world, plan.row_offset)
    if elast:
        # BASELINE.json configs[2]: the 3-D extension of examples/elasticity.py (lambda = 1, mu = .5/nu - 1 with nu = .3, body force -e_z)
        Ds = [engine.form_elasticity(3, 1., .5 / .3 - 1.)]
        Cs = [numpy.array([[0., 0, 0, 0], [0, 0, 0, 0], [-1., 0, 0, 0]])]
    else:
        Ds = [engine.form_stiffness(3), engine.form_mass(3)]
        Cs = [engine.form_load(3)]
    dev = torch.device('cuda', local)
    mats = [torch.empty(layout.nvalues, dtype=torch.float64, device=dev) for _ in Ds]
    vecs = [torch.empty(layout.nrows, dtype=torch.float64, device=dev) for _ in Cs]
    # window pointers: global slot s lives at window[s - off_lo]
    mptr = [m.data_ptr() - 8 * layout.off_lo for m in mats]
    vptr = [v.data_ptr() - 8 * layout.row_lo for v in vecs]

    def step():
        if rows_path:
            plan.assemble_rows_device(Ds, Cs, mptr, vptr, plane_range=layout.plane_range)
        else:
            for t in mats + vecs:
                t.zero_()
            plan.assemble_device(Ds, Cs, mptr, vptr, elem_range=layout.elem_range)
            if world > 1:
                distributed.exchange_interfaces(layout, mats, vecs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ctx.kernel_time()  # reset
    ctx.set_option('time_kernels', 1)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms, kernel_launches = ctx.kernel_time()
    ctx.set_option('time_kernels', 0)
    launches = ctx.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kernel_ms = t.tolist()
    ms_per_step = ms / args.steps
    ndofs_global = plan.ndofs
    value = ndofs_global / (ms_per_step * 1e-3)

    # sanity inside the bench: partition of unity on the assembled result of the last step (sum M == sum f)
    # (N>1: the ranks' windows tile the global arrays, so the global sums are the all-reduced window sums)
    sums = torch.stack([mats[-1].sum(), vecs[0].sum()])
    if world > 1:
        if not rows_path:  # shared planes are complete on both neighbours after the exchange: count the owned rows only
            sums = torch.stack([mats[1][:layout.off_own_hi - layout.off_lo].sum() if hasattr(layout, 'off_own_hi') else mats[1].sum(), vecs[0][layout.own_rows].sum()])
        dist.all_reduce(sums)
    msum, fsum = (float(x) for x in sums.tolist())
    if (world > 1 and not rows_path) or elast:
        msum = fsum = None  # the value windows of the slab layout overlap; the row-sum check is only meaningful for the rows path

    # roofline of the assembly kernel (rank-local bytes / rank-local kernel time)
    nnodes_local = (layout.elem_layers[1] - layout.elem_layers[0] if rows_path else (layout.elem_range[1] - layout.elem_range[0]) // (n * n)) + 1
    alg_bytes = 8. * (len(Ds) * layout.nvalues + len(Cs) * layout.nrows + 3 * nnodes_local * (n + 1) * (n + 1))
    # the dominant kernel's device time per step (a vector-valued step is one launch per row component)
    kernel_avg_ms = kernel_ms / max(args.steps, 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get('hbm_gbs', 6650.))
    achieved = alg_bytes / (kernel_avg_ms * 1e-3) / 1e9
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at the default workload
    if rows_path and world == 1 and n == 128 and p == 2:
        try:
            traffic = float(json.load(open(os.path.join(ROOT, 'profiles', 'r01', 'traffic.json')))['traffic_bytes'])
        except (OSError, ValueError, KeyError):
            pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (of fallback)',
                'kernel': 'k_rows3d (owner-computes assembly kernel; the only kernel of the step)' if rows_path else 'assembly kernel (zero-fill excluded)',
                'secondary_ceiling': 'FP64 pipe: the kernel issues ~1.2e9 warp-level FP64 instructions at 128^3 p=2 (see DESIGN.md), 2.1 ms at the measured 34 TFLOP/s DFMA rate', 'kernel_ms': kernel_avg_ms, 'algorithmic_bytes': alg_bytes,
                'kernel_share_of_step': kernel_avg_ms / ms_per_step, 'kernel_launches_per_step': kernel_launches / max(args.steps, 1)}

    # end-to-end through the host-buffer C-ABI call (single GPU path; ranks run it on their own slab problem)
    e2e = None
    if not args.no_e2e:
        lplan = plan if world == 1 else engine.Plan(ctx, [bspline.spline_basis_1d(n, p) for _ in range(3)], rules, make_nodes((n, n, n), seed=rank), ncomp=ncomp)
        lnodes = ctx.host_empty(lplan.nodes.shape)
        lnodes[...] = lplan.nodes
        hv = [ctx.host_empty(lplan.nnz) for _ in Ds]
        hr = [ctx.host_empty(lplan.ndofs) for _ in Cs]

        def e2e_step():
            lplan.update_nodes(lnodes)
            lplan.assemble_host(Ds, Cs, out_values=hv, out_rhs=hr)
            return hr[0][0]
        for _ in range(min(args.warmup, 3)):
            e2e_step()
        barrier()
        ksteps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e = {'value': lplan.ndofs * world / dt, 'unit': UNIT, 'h2d_bytes_per_step': int(lnodes.nbytes),
               'd2h_bytes_per_step': int(sum(a.nbytes for a in hv + hr)), 'ms_per_step': dt * 1e3, 'steps': ksteps,
               'note': 'b2_assemble_host with pinned host buffers; per rank an independent {}^3 problem'.format(n) if world > 1 else 'b2_assemble_host with pinned host buffers'}
        if world == 1 and not elast:
            # the host result of the e2e path doubles as a correctness check of the timed configuration
            assert abs(hv[1].sum() - hr[0].sum()) <= 1e-10 * abs(hr[0].sum()), 'sum(M) != sum(f)'

    # what follows the assembly when the matrix stays in HBM: y = K x on the analytic pattern
    after = None
    if world == 1 and not elast and rows_path:
        x = torch.rand(plan.ndofs, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        for _ in range(3):
            plan.spmv_device(mats[0], x, y)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(20):
            plan.spmv_device(mats[0], x, y)
        s1.record()
        torch.cuda.synchronize()
        sp_ms = s0.elapsed_time(s1) / 20
        sp_bytes = 8. * plan.nnz + 16. * plan.ndofs
        after = {'kernel': 'k_spmv (y = K x, analytic pattern, no column indices read)', 'ms': sp_ms, 'algorithmic_bytes': sp_bytes,
                 'achieved_GBps': sp_bytes / sp_ms * 1e-6, 'frac_of_hbm_peak': sp_bytes / sp_ms * 1e-6 / peak}

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu_n = args.cpu_n if not elast else min(args.cpu_n, 24)
        rate, dt, ndofs_s, cores = cpu_assembly_rate(cpu_n, p, elast=elast)
        cpu = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'seconds': dt,
               'sample': '{}^3 elements ({} dofs) of the same workload; C port of the reference algorithm (oracle/fem_oracle.c): threaded element loop + serial stable sort/unique/accumulate'.format(cpu_n, ndofs_s)}

    if rank == 0:
        line = {
            'metric': (METRIC_ELAST if elast else METRIC).replace('p=2', 'p={}'.format(p)), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(n, p, world, elast), 'ndofs': ndofs_global, 'nnz_per_matrix': plan.nnz, 'nelems': plan.ntotal,
                       'step': ('one owner-computes assembly launch (every value written once; no zero-fill, no exchange)' if rows_path else
                                'zero K,M,f + one assembly launch' + (' + neighbour exchange of shared dof planes (NCCL send/recv)' if world > 1 else '')),
                       'l2': 'outputs {:.2f} GB per rank >> 126 MB L2 (no flush needed)'.format(8e-9 * (len(Ds) * layout.nvalues + layout.nrows)),
                       'parallelism': ('dof planes along x owned per rank, overlap layers re-integrated, no collective' if rows_path else 'element slabs along x, one rank per GPU') if world > 1 else 'single GPU'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clk,
        }
        if after is not None:
            line['device_matrix'] = after
        if msum is not None:
            line['config']['check_sumM_minus_sumf'] = msum - fsum
            assert abs(msum - fsum) <= 1e-10 * abs(fsum), 'partition of unity violated: sum(M) != sum(f)'
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n', '--nelems', dest='n', type=int, default=128, help='elements per direction and GPU (under torchrun spell it --nelems: the launcher claims the abbreviation --n)')
    ap.add_argument('--degree', type=int, default=2)
    ap.add_argument('--cpu-n', type=int, default=40, help='elements per direction of the bounded CPU sample')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--workload', default='poisson', choices=['poisson', 'elasticity'], help='poisson: BASELINE.json configs[1] (the metric); elasticity: configs[2] (use --n 96)')
    ap.add_argument('--path', default='rows', choices=['rows', 'scatter'], help='rows: owner-computes kernel (default); scatter: element-scatter kernels')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
