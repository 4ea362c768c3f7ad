set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
nproc > gpurun_out/r2h_nproc.txt; free -g >> gpurun_out/r2h_nproc.txt; lscpu | head -20 >> gpurun_out/r2h_nproc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err
timeout 300 python bench.py --workload file:cfg2 --steps 3 > gpurun_out/r2h_file_cfg2.json 2> gpurun_out/r2h_file_cfg2.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "in_chain" 2>&1 | tail -15 > gpurun_out/r2h_pytest.log
ls -la gpurun_out | tail -5
