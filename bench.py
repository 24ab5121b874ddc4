#!/usr/bin/env python
'''Benchmark of the element-integration hot path (BASELINE.json: assembled DOFs/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload poisson|elasticity|nurbs_p4|fcm] [--n 128] [--degree 2] [--scaling weak|strong] [--path rows|scatter]

Workloads (BASELINE.json configs[1..4]); a "step" is one pass of the hot path over the whole problem, every input resident in HBM:

  poisson     configs[1] (THE metric): 3-D Poisson on n^3 elements (128^3), degree-p B-splines, Gauss degree 2p, K + M + f in one pass
              on the GENERAL-geometry code path (perturbed nodes: per-point Jacobians, no uniform-mesh shortcut).
  elasticity  configs[2]: 3-D linear elasticity (use --n 96), 3 components, K + f.
  nurbs_p4    configs[3]: examples/platewithhole.py in NURBS mode refined to 128 x 256 elements, degree-4 rational basis,
              plane-strain elasticity K + f, Gauss degree 8; element-set kernel, block products on the FP64 tensor cores (DMMA).
  fcm         configs[4]: finite-cell Poisson on a ball cut out of n^3 (64^3), ragged octree quadrature (depth 2), K + M + f.

N > 1 (torchrun): --scaling weak (default) grows the mesh with N ((N n) x n x n; rank r owns 1/N of the dof planes and
re-integrates the p element layers below its first plane: no collective); --scaling strong keeps the mesh fixed
((2n) x n x n for poisson) and divides it; --path scatter selects element slabs + ONE NCCL neighbour exchange of the p
shared dof planes (the north-star collective) instead.

ONE JSON line on rank 0.  `value`: device-timed (CUDA events, max over ranks) through the public API where there is one
(N = 1 structured: Sample.integrate_device).  `e2e`: the same pass through the host-buffer C-ABI call, H2D and D2H inside the
timed region.  `parity`: the SAME workload at the CPU sample size assembled on the GPU and compared with the oracle
(pattern bit-exact, values <= 1e-12) -- asserted.  `roofline`: algorithmic bytes over the dominant kernel's own device time
against the measured HBM peak.  `cpu_baseline`: the UNMODIFIED reference (baseline/_ref, NUTILS_NPROCS = all cores) on a
bounded sample, beside the C port of its algorithm.  --impl reference prints the reference's own line.
'''

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REFDIR = os.path.join(ROOT, 'baseline', '_ref')

UNIT = 'DOF/s'
METRICS = {
    'poisson': 'assembled DOFs/sec (stiffness+mass+RHS), 3D Poisson p={p}',
    'elasticity': 'assembled DOFs/sec (stiffness+RHS), 3D elasticity p={p}',
    'nurbs_p4': 'assembled DOFs/sec (stiffness+RHS), 2D NURBS plate-with-hole elasticity p={p}',
    'fcm': 'assembled DOFs/sec (stiffness+mass+RHS), 3D finite-cell Poisson p={p}',
}
LAMBDA, MU = 1., .5 / .3 - 1.   # configs[2]: lambda = 1, mu = .5/nu - 1 with nu = .3 (examples/elasticity.py)


def make_nodes(shape, seed=0, warp=.2):
    'perturbed nodal coordinates of the unit-spaced grid, float64[3, ...]; deterministic'
    rng = numpy.random.RandomState(seed)
    verts = [numpy.linspace(0, n / shape[-1], n + 1) for n in shape]
    X = numpy.stack(numpy.meshgrid(*verts, indexing='ij'))
    X += warp / shape[-1] * (rng.rand(*X.shape) - .5)
    return X


def source_digest():
    'identifies the kernel sources a committed ncu figure belongs to'
    h = hashlib.sha256()
    for name in sorted(os.listdir(os.path.join(ROOT, 'nutils_b200', 'csrc'))):
        if name.endswith(('.cu', '.cuh')):
            h.update(open(os.path.join(ROOT, 'nutils_b200', 'csrc', name), 'rb').read())
    return h.hexdigest()[:16]


def ncu_figures(key):
    '''dram traffic / tensor-pipe utilisation of the dominant kernel from the committed ncu capture of THIS workload
    (profiles/r02/ncu_metrics.json), only if the capture was taken from the kernel sources that are being run'''
    try:
        rec = json.load(open(os.path.join(ROOT, 'profiles', 'r02', 'ncu_metrics.json')))[key]
    except (OSError, ValueError, KeyError):
        return None
    if rec.get('source_digest') != source_digest():
        return {'stale': 'committed ncu capture belongs to other kernel sources ({})'.format(rec.get('source_digest'))}
    return rec


def relerr(a, b):
    n = numpy.linalg.norm(b)
    return float(numpy.linalg.norm(numpy.asarray(a) - numpy.asarray(b)) / (n if n else 1.))


def rowsum_relerr(va, vb, rowptr):
    starts = numpy.asarray(rowptr[:-1])
    nonempty = numpy.diff(rowptr) > 0
    ra = numpy.add.reduceat(va, starts[nonempty])
    rb = numpy.add.reduceat(vb, starts[nonempty])
    return float(abs(ra - rb).max() / numpy.add.reduceat(abs(vb), starts[nonempty]).max())


def parity_record(mine, ref, pattern_mine, pattern_ref, rhs_mine, rhs_ref, sample):
    rec = {'sample': sample,
           'pattern_equal': bool(numpy.array_equal(pattern_mine[0], pattern_ref[0]) and numpy.array_equal(pattern_mine[1], pattern_ref[1])),
           'frob_rel': max(relerr(a, b) for a, b in zip(mine, ref)),
           'rowsum_rel': max(rowsum_relerr(a, b, pattern_ref[0]) for a, b in zip(mine, ref)),
           'rhs_rel': max(relerr(a, b) for a, b in zip(rhs_mine, rhs_ref)), 'tol': 1e-12}
    rec['ok'] = bool(rec['pattern_equal'] and rec['frob_rel'] <= 1e-12 and rec['rowsum_rel'] <= 1e-12 and rec['rhs_rel'] <= 1e-12)
    return rec


# ---- CPU arms --------------------------------------------------------------------------------------------------------------

def structured_forms(workload):
    from nutils_b200 import engine
    if workload == 'elasticity':
        return [engine.form_elasticity(3, LAMBDA, MU)], [numpy.array([[0., 0, 0, 0], [0, 0, 0, 0], [-1., 0, 0, 0]])]
    return [engine.form_stiffness(3), engine.form_mass(3)], [engine.form_load(3)]


def port_problem(workload, n, degree):
    from nutils_b200 import bspline, points
    from oracle import fem_oracle
    b1 = [bspline.spline_basis_1d(n, degree) for _ in range(3)]
    rules = points.tensor_gauss(3, 2 * degree)
    return fem_oracle.Problem((n,) * 3, [degree] * 3, [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                              [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], make_nodes((n,) * 3), ncomp=3 if workload == 'elasticity' else 1), b1, rules


def port_assemble(workload, n, degree, nthreads=0):
    '''the C restatement of the reference algorithm (oracle/fem_oracle.c: threaded element loop + serial stable sort /
    unique / accumulate) on an n^3 sample of the workload: (DOF/s, seconds, ndofs, threads, matrices, vectors, tables)'''
    from oracle import c_oracle
    prob, b1, rules = port_problem(workload, n, degree)
    Ds, Cs = structured_forms(workload)
    c_oracle.lib()
    t0 = time.perf_counter()
    mats, vecs = c_oracle.assemble(prob, [('generic', D) for D in Ds], [('generic', C) for C in Cs], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return prob.ndofs / dt, dt, prob.ndofs, (nthreads or c_oracle.max_threads()), mats, vecs, (prob, b1, rules)


def have_reference():
    return os.path.isdir(os.path.join(REFDIR, 'nutils'))


def run_refworker(args):
    '''(subprocess, no CUDA in this process) time the UNMODIFIED reference on a bounded sample of the workload:
    f = evaluable.compile(as_csr(K), ..., F); f({}) -- compile excluded, first call timed (constants are cached afterwards,
    evaluable.py:6790-6821); NUTILS_NPROCS of the environment selects the fork-parallel loop (parallel.py:128-154).'''
    sys.path.insert(0, REFDIR)
    os.environ.setdefault('NUTILS_MATRIX', 'scipy')
    from nutils import mesh, function, evaluable
    from nutils.expression_v2 import Namespace
    p = args.degree

    def build(n):
        if args.workload in ('poisson', 'elasticity'):
            verts = [numpy.linspace(0, 1, n + 1)] * 3
            topo, geom0 = mesh.rectilinear(verts)
            X = make_nodes((n,) * 3)
            geom = (topo.basis('spline', degree=1) * X.reshape(3, -1)).sum(-1)   # mesh.py:55-57 with the perturbed nodes
            if args.workload == 'poisson':
                basis = topo.basis('spline', degree=p)
                g, J = basis.grad(geom), function.J(geom)
                ints = [topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=2 * p), topo.integral(basis[:, None] * basis[None, :] * J, degree=2 * p),
                        topo.integral(basis * J, degree=2 * p)]
                return [evaluable.as_csr(ints[0].as_evaluable_array), evaluable.as_csr(ints[1].as_evaluable_array), ints[2].as_evaluable_array], len(basis)
            ns = Namespace()
            ns.δ = function.eye(3)
            ns.x = geom
            ns.define_for('x', gradient='∇', jacobians=('dV',))
            ns.u = topo.field('u', btype='spline', degree=p, shape=[3])
            ns.v = topo.field('v', btype='spline', degree=p, shape=[3])
            ns.λ, ns.μ = LAMBDA, MU
            ns.ε_ij = '.5 (∇_i(u_j) + ∇_j(u_i))'
            ns.σ_ij = 'λ ε_kk δ_ij + 2 μ ε_ij'
            ns.q_i = '-δ_i2'
            res = topo.integral('(∇_j(v_i) σ_ij - v_i q_i) dV' @ ns, degree=2 * p).as_evaluable_array
            argv = [a for a in res.arguments if getattr(a, 'name', None) == 'v'][0]
            argu = [a for a in res.arguments if getattr(a, 'name', None) == 'u'][0]
            r = evaluable._flat(evaluable.derivative(res, argv))
            jac = evaluable._flat(evaluable.derivative(r, argu).simplified, 2)
            r0 = evaluable.replace_arguments(r, {'u': evaluable.zeros_like(argu)}).simplified
            return [evaluable.as_csr(jac), r0], 3 * (n + p) ** 3
        if args.workload == 'nurbs_p4':
            from nutils.solver import System
            topo, geom0 = mesh.rectilinear([1, 2])
            bs = topo.basis('spline', degree=2)
            cw = numpy.ones(12)
            cw[1:3] = .5 + .25 * numpy.sqrt(2)
            wf = bs @ cw
            radius = .5
            A = 0, 0, 0
            B = (2**.5 - 1) * radius, .3 * (radius + 1) / 2, 1
            C = radius, (radius + 1) / 2, 1
            geom = (bs * cw / wf) @ numpy.array([[A, B, C, C], [C, C, B, A]]).T.reshape(-1, 2)
            topo = topo.refine(n)
            ab = topo.basis('spline', degree=p)
            sqr = topo.integral((function.field('w', ab) - wf)**2, degree=9)
            basis = ab * System(sqr, trial='w').solve()['w'] / wf
            ns = Namespace()
            ns.δ = function.eye(2)
            ns.x = geom
            ns.define_for('x', gradient='∇', jacobians=('dV',))
            ns.u = function.field('u', basis, shape=[2])
            ns.v = function.field('v', basis, shape=[2])
            ns.λ, ns.μ = .6, .7
            ns.ε_ij = '(∇_j(u_i) + ∇_i(u_j)) / 2'
            ns.σ_ij = 'λ ε_kk δ_ij + 2 μ ε_ij'
            res = topo.integral('(∇_j(v_i) σ_ij - v_1) dV' @ ns, degree=2 * p).as_evaluable_array
            argv = [a for a in res.arguments if getattr(a, 'name', None) == 'v'][0]
            argu = [a for a in res.arguments if getattr(a, 'name', None) == 'u'][0]
            r = evaluable._flat(evaluable.derivative(res, argv))
            jac = evaluable._flat(evaluable.derivative(r, argu).simplified, 2)
            r0 = evaluable.replace_arguments(r, {'u': evaluable.zeros_like(argu)}).simplified
            return [evaluable.as_csr(jac), r0], 2 * len(ab)
        # fcm: sphere in a box, adaptive cut-cell quadrature (topology.py:1604-1657); the trim is host work, excluded
        topo0, geom = mesh.rectilinear([numpy.linspace(-1, 1, n + 1)] * 3)
        topo = topo0.trim(.8 - numpy.linalg.norm(geom), maxrefine=2)
        basis = topo.basis('spline', degree=p)
        g, J = basis.grad(geom), function.J(geom)
        ints = [topo.integral((g[:, None, :] * g[None, :, :]).sum(-1) * J, degree=2 * p), topo.integral(basis[:, None] * basis[None, :] * J, degree=2 * p),
                topo.integral(basis * J, degree=2 * p)]
        return [evaluable.as_csr(ints[0].as_evaluable_array), evaluable.as_csr(ints[1].as_evaluable_array), ints[2].as_evaluable_array], len(basis)

    def once(funcs):
        f = evaluable.compile(tuple(funcs))
        t0 = time.perf_counter()
        f({})
        return time.perf_counter() - t0

    sizes = {'poisson': [6, 8, 12, 16, 20, 24, 28, 32, 40], 'elasticity': [4, 6, 8, 10, 12, 16, 20, 24], 'nurbs_p4': [1, 2, 3, 4, 5], 'fcm': [4, 6, 8, 10, 12, 16]}[args.workload]
    # calibrate on the smallest size, then take the largest size whose run fits the time budget (cost ~ number of elements)
    funcs, ndofs = build(sizes[0])
    t_small = once(funcs)
    nel = lambda s: (2 * 4 ** s) if args.workload == 'nurbs_p4' else s ** 3
    pick = sizes[0]
    for s in sizes:
        if 2. * t_small * nel(s) / nel(sizes[0]) * (args.steps + args.warmup) <= args.ref_budget:   # 2: the post-loop sort grows faster than the loop
            pick = s
    if args.ref_n:
        pick = args.ref_n
    funcs, ndofs = build(pick)
    times = [once(funcs) for _ in range(args.warmup + args.steps)][args.warmup:]
    print(json.dumps({'size': pick, 'ndofs': int(ndofs), 'seconds': times, 'nprocs': int(os.environ.get('NUTILS_NPROCS', 1))}))


def reference_sample(args, steps, warmup, budget):
    'run the refworker subprocess on all host cores; returns its record or None when baseline/_ref is absent'
    if not have_reference():
        return None
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    env = dict(os.environ, NUTILS_NPROCS=str(cores), NUTILS_MATRIX='scipy', OMP_NUM_THREADS='1', CUDA_VISIBLE_DEVICES='')
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', '_refworker', '--workload', args.workload, '--degree', str(args.degree), '--steps', str(steps),
           '--warmup', str(warmup), '--ref-budget', str(budget), '--ref-n', str(args.ref_n)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1800)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    if out.returncode or not lines:
        return {'error': (out.stderr or out.stdout)[-400:]}
    rec = json.loads(lines[-1])
    rec['cores'] = cores
    return rec


def sample_text(workload, rec):
    s = rec['size']
    what = {'poisson': '{0}^3 elements'.format(s), 'elasticity': '{0}^3 elements'.format(s), 'nurbs_p4': 'plate-with-hole refined {0}x ({1}x{2} elements)'.format(s, 2 ** s, 2 ** (s + 1)),
            'fcm': 'sphere trimmed out of {0}^3 elements, maxrefine=2'.format(s)}[workload]
    return '{} ({} dofs) of the workload; unmodified evalf/nutils from baseline/_ref, NUTILS_NPROCS={} (fork-parallel element loop, parallel.py:128-154), first call of the compiled evaluable'.format(what, rec['ndofs'], rec['nprocs'])


def workload_text(args, N=1):
    p = args.degree
    if args.workload == 'nurbs_p4':
        return '2D plate-with-hole NURBS (examples/platewithhole.py:51-86) refined {}x = {}x{} elements, degree-{} rational basis, 2 components, gauss{} ({} pts), K+f'.format(
            args.nrefine, 2 ** args.nrefine, 2 ** (args.nrefine + 1), p, 2 * p, (p + 1) ** 2)
    if args.workload == 'fcm':
        return '3D finite-cell Poisson, ball r=0.8 in [-1,1]^3 on {0}^3 elements, p={1} spline, octree quadrature depth {2} on cut cells (ragged), K+M+f'.format(args.n, p, args.fcm_depth)
    n = args.n
    shape = strong_shape(args) if args.scaling == 'strong' else (n * N, n, n)
    mesh = 'x'.join(map(str, shape))
    if args.workload == 'elasticity':
        return '3D linear elasticity {} p={} spline (3 components), gauss{} ({} pts), K+f, nodal multilinear geometry (per-point Jacobians)'.format(mesh, p, 2 * p, (p + 1) ** 3)
    return '3D Poisson {} p={} spline, gauss{} ({} pts), K+M+f, nodal multilinear geometry (per-point Jacobians)'.format(mesh, p, 2 * p, (p + 1) ** 3)


def strong_shape(args):
    return (2 * args.n, args.n, args.n) if args.workload == 'poisson' else (args.n,) * 3


def run_reference(args):
    'the reference arm: rank 0 times the reference itself on the host cores; the other ranks exit'
    if int(os.environ.get('RANK', 0)) != 0:
        return
    metric = METRICS[args.workload].format(p=args.degree)
    rec = reference_sample(args, args.steps, args.warmup, 150.)
    if rec is None or 'error' in rec:
        if args.workload not in ('poisson', 'elasticity'):
            print(json.dumps({'impl': 'reference', 'unavailable': 'baseline/_ref is not installed and the C port covers the structured workloads only'}))
            return
        # baseline/_ref is absent (scripts/install_reference.py was not run): the C port of the reference algorithm stands in
        n = 32 if args.warmup + args.steps <= 40 else 24
        times = []
        for i in range(args.warmup + args.steps):
            rate, dt, ndofs, cores, _, _, _ = port_assemble(args.workload, n if args.workload == 'poisson' else min(n, 20), args.degree)
            if i >= args.warmup:
                times.append(dt)
        kind, sample = 'port', '{}^3 elements ({} dofs) of the workload; C port of the reference algorithm (baseline/_ref not installed: {})'.format(n, ndofs, rec)
    else:
        times, ndofs, cores, kind, sample = rec['seconds'], rec['ndofs'], rec['cores'], 'reference', sample_text(args.workload, rec)
    t = sum(times) / len(times)
    value = ndofs / t
    print(json.dumps({
        'impl': 'reference', 'metric': metric, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_text(args, args.gpus), 'timed_sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))


# ---- host placement of a rank (e2e over PCIe) ----------------------------------------------------------------------------------

def numa_nodes():
    'cpu lists of the NUMA nodes of this host: {node: [cpus]}'
    out = {}
    base = '/sys/devices/system/node'
    try:
        for name in sorted(os.listdir(base)):
            if name.startswith('node') and name[4:].isdigit():
                cpus = []
                for part in open(os.path.join(base, name, 'cpulist')).read().strip().split(','):
                    if part:
                        a, _, b = part.partition('-')
                        cpus.extend(range(int(a), int(b or a) + 1))
                if cpus:
                    out[int(name[4:])] = cpus
    except OSError:
        pass
    return out


def place_rank(local, world, policy):
    '''Pin this rank's threads -- and with them the first-touch placement of its pinned host buffers -- to a NUMA node before
    anything is allocated.  policy 'gpu': the node the GPU's PCIe root reports (sysfs numa_node); 'spread': ranks round-robin over
    the nodes (all GPUs of these boxes report node 0, whose memory controllers then take the D2H traffic of every rank); 'none'.
    Returns a description for the JSON line.'''
    nodes = numa_nodes()
    allowed = set(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else set()
    if policy == 'none' or len(nodes) < 2 or not allowed:
        return {'policy': 'none', 'nodes': len(nodes)}
    node = None
    if policy == 'gpu':
        try:
            import torch
            pr = torch.cuda.get_device_properties(local)
            bdf = '{:04x}:{:02x}:{:02x}.0'.format(pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            node = int(open('/sys/bus/pci/devices/{}/numa_node'.format(bdf)).read())
        except Exception:
            node = None
        if node is None or node < 0:
            policy = 'spread'
    if policy == 'spread':
        keys = sorted(nodes)
        node = keys[local * len(keys) // max(world, 1)] if world > 1 else keys[0]
    cpus = sorted(set(nodes.get(node, [])) & allowed)
    if not cpus:
        return {'policy': 'none', 'nodes': len(nodes), 'note': 'no allowed cpu on node {}'.format(node)}
    os.sched_setaffinity(0, cpus)
    return {'policy': policy, 'node': node, 'cpus': len(cpus), 'nodes': len(nodes)}


# ---- clocks ----------------------------------------------------------------------------------------------------------------

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


# ---- GPU arm: workloads ----------------------------------------------------------------------------------------------------

class Structured:
    'configs[1] / configs[2]: the owner-computes rows kernel (k_geom3d + k_rows3d) or, with --path scatter, element slabs + one exchange'

    def __init__(self, args, ctx, rank, world, dev):
        import torch
        from nutils_b200 import bspline, points, engine, distributed, mesh, function
        self.args, self.ctx, self.rank, self.world = args, ctx, rank, world
        n, p = args.n, args.degree
        self.elast = args.workload == 'elasticity'
        self.ncomp = 3 if self.elast else 1
        self.shape = strong_shape(args) if args.scaling == 'strong' else (n * world, n, n)
        self.rows_path = args.path == 'rows'
        self.Ds, self.Cs = structured_forms(args.workload)
        nodes = make_nodes(self.shape)
        self.api = world == 1 and self.rows_path and not self.elast
        if self.api:
            # the public API: mesh.rectilinear -> basis -> integrands -> Sample.integrate_device (what a user script calls)
            topo, geom0 = mesh.rectilinear(list(self.shape))
            geom = topo.nodal_geometry(nodes)
            basis = topo.basis('spline', degree=p)
            g, J = basis.grad(geom), function.J(geom)
            self.funcs = [(g[:, None, :] * g[None, :, :]).sum(-1) * J, function.outer(basis) * J, basis * J]
            self.sample = topo.sample('gauss', 2 * p)
            self.out = self.sample.integrate_device_buffers(self.funcs)
            self.plan = self.sample.plan(basis, geom)
            self.nvalues, self.nrows, self.nlayers = self.plan.nnz, self.plan.ndofs, self.shape[0]
            self.step = lambda: self.sample.integrate_device(self.funcs, out=self.out)
            self.api_call = 'Sample.integrate_device([K, M, f], out=buffers) (nutils_b200/sample.py; the Nutils-style public API)'
        else:
            b1 = [bspline.spline_basis_1d(m, p) for m in self.shape]
            self.plan = plan = engine.Plan(ctx, b1, points.tensor_gauss(3, 2 * p), nodes, ncomp=self.ncomp)
            self.layout = lay = (distributed.PlaneLayout if self.rows_path else distributed.SlabLayout)(b1, self.ncomp, rank, world, plan.row_offset)
            self.mats = [torch.empty(lay.nvalues, dtype=torch.float64, device=dev) for _ in self.Ds]
            self.vecs = [torch.empty(lay.nrows, dtype=torch.float64, device=dev) for _ in self.Cs]
            mptr = [m.data_ptr() - 8 * lay.off_lo for m in self.mats]   # window pointers: global slot s lives at window[s - off_lo]
            vptr = [v.data_ptr() - 8 * lay.row_lo for v in self.vecs]
            self.nvalues, self.nrows = lay.nvalues, lay.nrows
            self.nlayers = (lay.elem_layers[1] - lay.elem_layers[0]) if self.rows_path else (lay.elem_range[1] - lay.elem_range[0]) // (self.shape[1] * self.shape[2])
            self.halo_bytes = 0
            if self.rows_path:
                self.step = lambda: plan.assemble_rows_device(self.Ds, self.Cs, mptr, vptr, plane_range=lay.plane_range)
            else:
                self.halo_bytes = 8 * sum((vs.stop - vs.start) * len(self.Ds) + (rs.stop - rs.start) * len(self.Cs) for peer, vs, rs in lay.neighbours)

                def step():
                    for t in self.mats + self.vecs:
                        t.zero_()
                    plan.assemble_device(self.Ds, self.Cs, mptr, vptr, elem_range=lay.elem_range)
                    if world > 1:
                        distributed.exchange_interfaces(lay, self.mats, self.vecs)
                self.step = step
            self.api_call = 'engine.Plan.assemble_rows_device (b2_assemble_rows_device) on the rank\'s window' if self.rows_path else 'zero + engine.Plan.assemble_device + distributed.exchange_interfaces'
        self.ndofs_global = self.plan.ndofs
        self.nnz = self.plan.nnz
        self.nelems = self.plan.ntotal
        # algorithmic bytes of this rank: 8 B per stored value and rhs entry + 24 B per node of the layers it integrates
        self.alg_bytes = 8. * (len(self.Ds) * self.nvalues + len(self.Cs) * self.nrows + 3 * (self.nlayers + 1) * (self.shape[1] + 1) * (self.shape[2] + 1))
        self.kernel = ('k_rows3d (owner-computes tile kernel; its geometry comes from k_geom3d by TMA)' if self.rows_path else 'k_assemble_scalar3d (element scatter; zero-fill and exchange excluded)')
        if self.elast and self.rows_path:
            self.kernel = ('k_rows3d, 6 of the 9 component blocks of the symmetric form: 3 diagonal blocks on the symmetric scalar pipeline (k_geom3d + TMA), '
                           '3 blocks above the diagonal on the general pipeline with transposed stores; figure = all launches of a step together')
        self.ncu_key = '{}_n{}_p{}'.format(args.workload, args.n, p) if (world == 1 and self.rows_path) else None

    def sums(self):
        'partition of unity on the assembled result: (sum M, sum f) of this rank, or None'
        import torch
        if self.elast or not self.rows_path:
            return None
        if self.api:
            M = self.out[1].values.to_host()
            return float(M.sum()), float(self.out[2].to_host().sum())
        return float(self.mats[1].sum()), float(self.vecs[0].sum())

    def e2e(self, barrier):
        'the host-buffer C-ABI call: H2D of the nodal coordinates, D2H of every result, pinned host memory'
        from nutils_b200 import bspline, points, engine
        args, ctx = self.args, self.ctx
        n, p = args.n, args.degree
        if self.world == 1 and not self.api:
            lplan = self.plan
        elif self.world == 1:
            lplan = self.plan
        else:
            lplan = engine.Plan(ctx, [bspline.spline_basis_1d(n, p) for _ in range(3)], points.tensor_gauss(3, 2 * p), make_nodes((n, n, n), seed=self.rank), ncomp=self.ncomp)
        lnodes = ctx.host_empty(lplan.nodes.shape)
        lnodes[...] = lplan.nodes
        hv = [ctx.host_empty(lplan.nnz) for _ in self.Ds]
        hr = [ctx.host_empty(lplan.ndofs) for _ in self.Cs]

        def e2e_step():
            lplan.update_nodes(lnodes)
            lplan.assemble_host(self.Ds, self.Cs, out_values=hv, out_rhs=hr)
            return hr[0][0]
        for _ in range(2):
            e2e_step()
        barrier()
        ksteps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        if self.world == 1 and not self.elast:
            assert abs(hv[1].sum() - hr[0].sum()) <= 1e-10 * abs(hr[0].sum()), 'sum(M) != sum(f)'
        note = 'b2_assemble_host with pinned host buffers' + ('; per rank an independent {}^3 problem'.format(n) if self.world > 1 else '')
        return dt, lplan.ndofs * self.world, int(lnodes.nbytes), int(sum(a.nbytes for a in hv + hr)), ksteps, note

    def parity(self):
        'the timed configuration at the CPU sample size: GPU (same entry point, same kernels) against the C oracle'
        from nutils_b200 import engine
        args = self.args
        cpu_n = args.cpu_n if not self.elast else min(args.cpu_n, 24)
        rate, dt, ndofs, cores, mats, vecs, (prob, b1, rules) = port_assemble(args.workload, cpu_n, args.degree)
        plan = engine.Plan(self.ctx, b1, rules, prob.nodes, ncomp=self.ncomp)
        vals, rhs = plan.assemble_host(self.Ds, self.Cs)
        rec = parity_record(vals, [m[0] for m in mats], plan.csr_pattern(), (mats[0][1], mats[0][2]), rhs, vecs,
                            '{}^3 elements ({} dofs), same forms, geometry generator and kernels as the timed workload, against oracle/fem_oracle.c'.format(cpu_n, ndofs))
        port = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'seconds': dt,
                'sample': '{}^3 elements ({} dofs); C port of the reference algorithm (oracle/fem_oracle.c): threaded element loop + serial stable sort/unique/accumulate'.format(cpu_n, ndofs)}
        return rec, port


class ElemSetWorkload:
    'configs[3] (NURBS p=4) and configs[4] (finite cell): the element-set kernel k_assemble_elemset (block products as DMMA) + zero-fill'

    def __init__(self, args, ctx, rank, world, dev):
        import torch
        from nutils_b200 import engine, distributed
        self.args, self.ctx, self.rank, self.world = args, ctx, rank, world
        self.tables = t = self.make_tables(args.nrefine if args.workload == 'nurbs_p4' else args.n)
        t0 = time.perf_counter()
        self.plan = plan = engine.ElemSetPlan(ctx, **t['plan'])
        ctx.synchronize()
        self.plan_seconds = time.perf_counter() - t0
        self.Ds, self.Cs = t['Ds'], t['Cs']
        self.ndofs_global, self.nnz, self.nelems = plan.ndofs, plan.nnz, plan.nsel
        if world > 1:
            kw = {k: t['plan'][k] for k in ('elem_ids', 'qoff', 'renumber', 'nbasis_new') if k in t['plan']}
            self.layout = lay = distributed.ElemSetLayout(t['plan']['bases'], plan.ncomp, rank, world, plan.row_offset, points_per_element=t.get('nq_uniform'), **kw)
            nvalues, nrows, off_lo, row_lo, sel = lay.nvalues, lay.nrows, lay.off_lo, lay.row_lo, lay.sel_range
            share = lay.cost_share
        else:
            nvalues, nrows, off_lo, row_lo, sel, share = plan.nnz, plan.ndofs, 0, 0, None, 1.
        self.mats = [torch.zeros(nvalues, dtype=torch.float64, device=dev) for _ in self.Ds]
        self.vecs = [torch.zeros(nrows, dtype=torch.float64, device=dev) for _ in self.Cs]
        mptr = [m.data_ptr() - 8 * off_lo for m in self.mats]
        vptr = [v.data_ptr() - 8 * row_lo for v in self.vecs]

        def step():
            for x in self.mats + self.vecs:
                x.zero_()
            plan.assemble_device(self.Ds, self.Cs, mptr, vptr, sel_range=sel)
            if world > 1:
                distributed.exchange_interfaces(lay, self.mats, self.vecs)
        self.step = step
        self.api_call = 'engine.ElemSetPlan.assemble_device (b2_assemble_elemset_device) after zeroing the outputs' + ('; one NCCL neighbour exchange' if world > 1 else '')
        self.nvalues, self.nrows = nvalues, nrows
        # 8 B per stored value / rhs entry + the per-element inputs: 8 (d+1) B per point (ragged) + control points
        self.alg_bytes = 8. * (len(self.Ds) * nvalues + len(self.Cs) * nrows) + share * t['input_bytes']
        self.kernel = 'k_assemble_elemset (one CTA per element; block products as mma.sync.m8n8k4.f64; fp64 RED scatter)'
        self.ncu_key = '{}_n{}_p{}'.format(args.workload, args.nrefine if args.workload == 'nurbs_p4' else args.n, args.degree) if world == 1 else None
        self.shape = t['shape']

    def make_tables(self, size):
        from nutils_b200 import engine, bspline, points, nurbs, fcm
        p = self.args.degree
        if self.args.workload == 'nurbs_p4':
            t = nurbs.plate_with_hole(size, p)
            C = numpy.zeros((2, 3))
            C[1, 0] = 1.
            npts = int(numpy.prod(t['shape'])) * (p + 1) ** 2
            return dict(plan=dict(bases=t['bases'], ncomp=2, rules=points.tensor_gauss(2, 2 * p), scale=t['scale'], rational=2, geom_spline=(t['gbases'], t['gctrl'], t['gweights'])),
                        Ds=[engine.form_elasticity(2, .6, .7)], Cs=[C], shape=t['shape'], nq_uniform=(p + 1) ** 2,
                        input_bytes=8. * (t['gctrl'].size + t['gweights'].size + t['scale'].size), npoints=npts)
        n = size
        elem_ids, qoff, qc, qw, ren, nbn = fcm.octree_ball(n, p, self.args.fcm_depth)
        b1 = [bspline.spline_basis_1d(n, p) for _ in range(3)]
        v = numpy.linspace(-1, 1, n + 1)
        nodes = numpy.stack(numpy.meshgrid(v, v, v, indexing='ij'))
        return dict(plan=dict(bases=b1, nodes=nodes, elem_ids=elem_ids, qoff=qoff, qcoords=qc, qweights=qw, renumber=ren, nbasis_new=nbn),
                    Ds=[engine.form_stiffness(3), engine.form_mass(3)], Cs=[engine.form_load(3)], shape=(n,) * 3,
                    input_bytes=8. * 4 * int(qoff[-1]) + 8. * len(elem_ids) + 24. * (n + 1) ** 3, npoints=int(qoff[-1]))

    def sums(self):
        if self.args.workload != 'fcm' or self.world > 1:
            return None
        return float(self.mats[1].sum()), float(self.vecs[0].sum())

    def e2e(self, barrier):
        '''host arrays in, host arrays out: every step uploads the tables of the element set (points, numbering, geometry), builds
        the pattern on the device, assembles and copies K, f back (engine.ElemSetPlan(...).assemble_host = b2_elemset_create +
        b2_pattern_create_elemset + b2_assemble_elemset_host)'''
        from nutils_b200 import engine
        t = self.tables
        h2d = sum(numpy.asarray(v).nbytes for v in t['plan'].values() if isinstance(v, numpy.ndarray))
        if 'geom_spline' in t['plan']:
            h2d += sum(numpy.asarray(a).nbytes for a in t['plan']['geom_spline'][1:])

        def e2e_step():
            plan = engine.ElemSetPlan(self.ctx, **t['plan'])
            return plan.assemble_host(self.Ds, self.Cs)
        e2e_step()
        barrier()
        ksteps = max(1, min(self.args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            vals, rhs = e2e_step()
        dt = (time.perf_counter() - t0) / ksteps
        d2h = int(sum(a.nbytes for a in vals + rhs))
        return dt, self.plan.ndofs * self.world, int(h2d), d2h, ksteps, 'engine.ElemSetPlan(host tables) + assemble_host per step: table upload, device pattern build, assembly, D2H' + ('; per rank the whole problem' if self.world > 1 else '')

    def parity(self):
        'a small instance of the same construction against the numpy oracle (oracle/fem_oracle.py)'
        from nutils_b200 import engine, points
        from oracle import fem_oracle
        p = self.args.degree
        small = 2 if self.args.workload == 'nurbs_p4' else 6
        t = self.make_tables(small)
        plan = engine.ElemSetPlan(self.ctx, **t['plan'])
        vals, rhs = plan.assemble_host(t['Ds'], t['Cs'])
        pl = t['plan']
        b1 = pl['bases']
        nd = len(b1)
        rules = points.tensor_gauss(nd, 2 * p)
        kw = dict(ncomp=pl.get('ncomp', 1))
        for key in ('elem_ids', 'qoff', 'qcoords', 'qweights', 'renumber', 'nbasis_new', 'scale', 'rational'):
            if key in pl:
                kw[key] = pl[key]
        nodes = pl.get('nodes')
        if 'geom_spline' in pl:
            gb, gctrl, gw = pl['geom_spline']
            kw['geom_spline'] = dict(degree=[b.degree for b in gb], coeffs=[b.coeffs for b in gb], setidx=[b.setidx for b in gb], start=[b.start for b in gb],
                                     ndofs_d=[b.ndofs for b in gb], ctrl=gctrl, weights=gw)
            nodes = numpy.zeros((nd,) + (0,) * nd)
        prob = fem_oracle.Problem(tuple(b.nelems for b in b1), [b.degree for b in b1], [b.coeffs for b in b1], [b.setidx for b in b1], [b.start for b in b1],
                                  [b.ndofs for b in b1], [r[0] for r in rules], [r[1] for r in rules], nodes, **kw)
        mats, vecs = fem_oracle.assemble(prob, [('generic', D) for D in t['Ds']], [('generic', C) for C in t['Cs']])
        rec = parity_record(vals, [m[0] for m in mats], plan.csr_pattern(), (mats[0][1], mats[0][2]), rhs, vecs,
                            '{} ({} dofs, {} points), same construction and kernels as the timed workload, against oracle/fem_oracle.py'.format(
                                'x'.join(map(str, t['shape'])) + ' elements', plan.ndofs, t['npoints']))
        return rec, None


def run_b200(args):
    import torch
    import torch.distributed as dist
    from nutils_b200 import engine

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit('launch with torchrun --nproc-per-node {} for --gpus {}'.format(args.gpus, args.gpus))
    placement = place_rank(local, world, args.numa if world > 1 else 'none')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    ctx = engine.Context.get(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    p = args.degree
    W = (Structured if args.workload in ('poisson', 'elasticity') else ElemSetWorkload)(args, ctx, rank, world, dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        W.step()
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
        W.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kernel_ms, kernel_launches = ctx.kernel_time()
    ctx.set_option('time_kernels', 0)
    launches = ctx.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None
    # the dominant kernel of the step = its longest launch (geometry + tile kernel, or zero-fill + assembly): one more step with the
    # context recording the maximum single-launch time instead of the sum
    dominant_ms = dominant_kernel_time(ctx, W, torch) if kernel_launches > args.steps else kernel_ms / max(args.steps, 1)
    if getattr(W, 'elast', False):
        dominant_ms = kernel_ms / max(args.steps, 1)   # vector-valued: the launches of the tile kernel per component block each write a part of the rows: taken together
    if world > 1:
        t = torch.tensor([ms, kernel_ms, dominant_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kernel_ms, dominant_ms = t.tolist()
    ms_per_step = ms / args.steps
    value = W.ndofs_global / (ms_per_step * 1e-3)

    sums = W.sums()
    check = None
    if sums is not None:
        s = torch.tensor(sums, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(s)
        check = float(s[0] - s[1])
        assert abs(check) <= 1e-10 * abs(float(s[1])), 'partition of unity violated: sum(M) != sum(f)'

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get('hbm_gbs', 6650.))
    kernel_avg_ms = kernel_ms / max(args.steps, 1)
    achieved = W.alg_bytes / (dominant_ms * 1e-3) / 1e9
    ncu = ncu_figures(W.ncu_key) if W.ncu_key else None
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': (ncu or {}).get('dram_bytes'), 'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (of fallback)',
                'kernel': W.kernel, 'kernel_ms': dominant_ms, 'all_kernels_ms_per_step': kernel_avg_ms, 'algorithmic_bytes': W.alg_bytes,
                'kernel_share_of_step': dominant_ms / ms_per_step, 'kernel_launches_per_step': kernel_launches / max(args.steps, 1),
                'frac_all_kernels': W.alg_bytes / (kernel_avg_ms * 1e-3) / 1e9 / peak,
                'secondary_ceiling': 'FP64 pipe (DFMA / DMMA): the binding unit of this path, see DESIGN.md section 4'}
    if ncu:
        roofline['ncu'] = ncu

    e2e = None
    if not args.no_e2e:
        dt, ndofs_e2e, h2d, d2h, ksteps, note = W.e2e(barrier)
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e = {'value': ndofs_e2e / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': dt * 1e3, 'steps': ksteps, 'note': note}

    after = None
    if world == 1 and args.workload == 'poisson' and args.path == 'rows':
        # what follows the assembly when the matrix stays in HBM: y = K x on the analytic pattern
        Kvals = W.out[0].values if W.api else W.mats[0]
        x = torch.rand(W.plan.ndofs, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        for _ in range(3):
            W.plan.spmv_device(Kvals, x, y)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(20):
            W.plan.spmv_device(Kvals, x, y)
        s1.record()
        torch.cuda.synchronize()
        sp_ms = s0.elapsed_time(s1) / 20
        sp_bytes = 8. * W.plan.nnz + 16. * W.plan.ndofs
        after = {'kernel': 'k_spmv (y = K x, analytic pattern, no column indices read)', 'ms': sp_ms, 'algorithmic_bytes': sp_bytes,
                 'achieved_GBps': sp_bytes / sp_ms * 1e-6, 'frac_of_hbm_peak': sp_bytes / sp_ms * 1e-6 / peak}

    parity = cpu = None
    if rank == 0 and not args.no_cpu:
        parity, port = W.parity()
        assert parity['ok'], 'GPU result differs from the oracle: {}'.format(parity)
        ref = reference_sample(args, 1, 0, 25.)
        if ref is not None and 'error' not in ref:
            tref = sum(ref['seconds']) / len(ref['seconds'])
            cpu = {'value': ref['ndofs'] / tref, 'unit': UNIT, 'cores': ref['cores'], 'kind': 'reference', 'seconds': tref, 'sample': sample_text(args.workload, ref)}
            if port:
                cpu['port'] = port
        elif port:
            cpu = dict(port, note='baseline/_ref not usable here: {}'.format(ref))

    if rank == 0:
        config = {'workload': workload_text(args, world), 'ndofs': W.ndofs_global, 'nnz_per_matrix': W.nnz, 'nelems': W.nelems,
                  'step': W.api_call, 'l2': 'outputs {:.2f} GB per rank >> 126 MB L2 (no flush needed)'.format(8e-9 * (len(W.Ds) * W.nvalues + W.nrows)),
                  'parallelism': 'single GPU' if world == 1 else ('{} scaling over {} ranks; '.format(args.scaling, world) + (
                      'dof planes along x owned per rank, overlap layers re-integrated, no collective' if getattr(W, 'rows_path', False) else
                      'element ranges per rank, ONE neighbour exchange (NCCL send/recv) of the shared dof rows'))}
        if getattr(W, 'halo_bytes', 0):
            config['halo_bytes_per_rank'] = W.halo_bytes
        if world > 1:
            config['host_placement_rank0'] = placement
        if check is not None:
            config['check_sumM_minus_sumf'] = check
        if hasattr(W, 'plan_seconds'):
            config['plan_seconds (table upload + device pattern build, excluded from the metric)'] = W.plan_seconds
        line = {'metric': METRICS[args.workload].format(p=p), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': config, 'roofline': roofline, 'parity': parity, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clk}
        if after is not None:
            line['device_matrix'] = after
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dominant_kernel_time(ctx, W, torch):
    'device time of the LONGEST kernel launch of a step (context option time_kernels = 2: b2_ctx_kernel_time returns the maximum); median of 5 steps'
    out = []
    for _ in range(5):
        ctx.kernel_time()
        ctx.set_option('time_kernels', 2)
        W.step()
        torch.cuda.synchronize()
        ms, n = ctx.kernel_time()
        ctx.set_option('time_kernels', 0)
        out.append(ms)
    return float(numpy.median(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', '_refworker'])
    ap.add_argument('--workload', default='poisson', choices=['poisson', 'elasticity', 'nurbs_p4', 'fcm'])
    ap.add_argument('--n', '--nelems', dest='n', type=int, default=None, help='elements per direction (and GPU under weak scaling); default 128 (poisson), 96 (elasticity), 64 (fcm); under torchrun spell it --nelems')
    ap.add_argument('--degree', type=int, default=None)
    ap.add_argument('--nrefine', type=int, default=7, help='nurbs_p4: refinement levels of the 1x2 patch (7: 128x256 elements)')
    ap.add_argument('--fcm-depth', type=int, default=2)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--path', default='rows', choices=['rows', 'scatter'], help='structured workloads: rows = owner-computes kernel (default); scatter = element-scatter kernels + exchange')
    ap.add_argument('--numa', default='spread', choices=['none', 'gpu', 'spread'], help='N > 1: NUMA placement of a rank (and of its pinned host buffers)')
    ap.add_argument('--cpu-n', type=int, default=40, help='elements per direction of the parity / C-port sample')
    ap.add_argument('--ref-n', type=int, default=0, help='size of the reference sample (0: fitted to the time budget)')
    ap.add_argument('--ref-budget', type=float, default=150.)
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.n is None:
        args.n = {'poisson': 128, 'elasticity': 96, 'fcm': 64, 'nurbs_p4': 128}[args.workload]
    if args.degree is None:
        args.degree = 4 if args.workload == 'nurbs_p4' else 2
    if args.impl == '_refworker':
        return run_refworker(args)
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
