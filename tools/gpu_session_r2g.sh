set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py tests/test_shard.py -m gpu -q --timeout 600 -k "iq_optim or dropin or shard" 2>&1 | tail -30 > gpurun_out/r2g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg5 --sharded-capture= > gpurun_out/r2g_bench_n2_cfg5.json 2> gpurun_out/r2g_bench_n2_cfg5.err
ls -la gpurun_out | tail -5
