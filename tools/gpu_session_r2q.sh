set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n4.json 2> gpurun_out/r2q_bench_n4.err
nvidia-smi topo -m > gpurun_out/r2q_topo.txt 2>&1
ls -la gpurun_out | tail -3
