set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# the driver's scaling run, back to back on one 8-GPU box: N = 1, 2, 4, 8 with the default flags
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2q_bench_n$n.json 2> gpurun_out/r2q_bench_n$n.err
done
nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1
ls -la gpurun_out | tail -3
