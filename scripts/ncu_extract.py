#!/usr/bin/env python
'''Reduce an ncu report to the figures bench.py quotes: writes / updates profiles/r02/ncu_metrics.json.

    python scripts/ncu_extract.py report.ncu-rep KEY "workload description" [kernel-substring]

KEY is the workload key of bench.py (e.g. poisson_n128_p2); the entry is stamped with the digest of the kernel sources
(bench.source_digest) so that bench.py only quotes it for the code it was measured on.  Runs where ncu is installed and the
report is readable (the build container), not on the GPU box.'''
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rep, key, workload = sys.argv[1:4]
want = sys.argv[4] if len(sys.argv) > 4 else ''
out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout  # a report, or its exported raw page
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {name: i for i, name in enumerate(hdr)}
sel = [r for r in rows[2:] if want in r[col['Kernel Name']]]
r = sel[-1]


def get(name, scale=1.):
    for k, i in col.items():
        if k == name or k.endswith(name):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                return None
            u = units[i]
            mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1., 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1., 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.}.get(u, 1.)
            return v * mult * scale
    return None


rec = {
    'workload': workload, 'kernel': r[col['Kernel Name']].strip(), 'report': os.path.basename(rep), 'source_digest': bench.source_digest(),
    'duration_ms_under_ncu': get('gpu__time_duration.sum', 1e3),
    'dram_bytes': (get('dram__bytes_read.sum') or 0.) + (get('dram__bytes_write.sum') or 0.),
    'dram_bytes_read': get('dram__bytes_read.sum'), 'dram_bytes_write': get('dram__bytes_write.sum'),
    'dram_throughput_pct': get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    'warp_instructions': get('smsp__inst_executed.sum'),
    'issue_slots_busy_pct': get('smsp__issue_active.avg.pct_of_peak_sustained_active') or get('sm__inst_issued.avg.pct_of_peak_sustained_active'),
    'fp64_pipe_pct': get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
    'tensor_pipe_pct': get('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'),
    'tensor_dmma_inst_pct_of_peak': get('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active'),
    'shared_pipe_pct': get('sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active'),
    'eligible_warps_per_scheduler': get('smsp__warps_eligible.avg.per_cycle_active'),
    'registers_per_thread': get('launch__registers_per_thread'),
}
path = os.path.join(ROOT, 'profiles', 'r02', 'ncu_metrics.json')
try:
    data = json.load(open(path))
except (OSError, ValueError):
    data = {}
data[key] = rec
json.dump(data, open(path, 'w'), indent=1, sort_keys=True)
print(json.dumps(rec, indent=1))
