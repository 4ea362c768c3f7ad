set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2u_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
IQGPU_DEBUG_AGC=1 timeout 300 python bench.py --workload cfg4 --steps 20 $B > gpurun_out/r2u_bench_cfg4.json 2> gpurun_out/r2u_bench_cfg4.err
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:agc_rms -s 2 -c 1 -o gpurun_out/r2u_agcrms_cfg4 python bench.py --workload cfg4 $N > gpurun_out/r2u_ncu_cfg4.log 2>&1
du -sh gpurun_out
