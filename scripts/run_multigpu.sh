#!/bin/bash
# Multi-GPU bench lines of one box (development helper; the driver's own scaling run is the judged one).
# usage: bash scripts/run_multigpu.sh MAXN   (MAXN = 8, 4 or 2: the GPUs of the box)
MAXN=${1:-8}
mkdir -p gpurun_out/mg
run() { # N tag extra-args
  N=$1; shift; TAG=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu "$@" > gpurun_out/mg/$TAG.json 2> gpurun_out/mg/$TAG.err
  tail -c 300 gpurun_out/mg/$TAG.err | tail -1
  cut -c1-180 gpurun_out/mg/$TAG.json
}
nvidia-smi topo -m | head -12 > gpurun_out/mg/topo_n$MAXN.txt
for N in 8 4 2; do
  if [ $N -le $MAXN ]; then
    E2E=--no-e2e; [ $N -eq $MAXN ] && E2E=
    run $N strong_poisson_n$N --scaling strong $E2E
    run $N strong_elasticity_n$N --workload elasticity --scaling strong $E2E
  fi
done
run $MAXN weak_poisson_n$MAXN
run $MAXN strong_poisson_scatter_n$MAXN --scaling strong --path scatter --no-e2e
