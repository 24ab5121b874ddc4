mkdir -p gpurun_out/mg
run() { N=$1; shift; TAG=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu "$@" > gpurun_out/mg/$TAG.json 2> gpurun_out/mg/$TAG.err
  cut -c1-180 gpurun_out/mg/$TAG.json; }
nvidia-smi topo -m | head -12 > gpurun_out/mg/topo_n8.txt
run 8 strong_poisson_n8 --scaling strong --no-e2e
run 8 strong_elasticity_n8 --workload elasticity --scaling strong --no-e2e
run 8 weak_poisson_n8
