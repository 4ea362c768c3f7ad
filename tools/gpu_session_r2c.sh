set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r2c_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for tc in 40 20 10 5; do
  IQGPU_DEBUG_AGC=1 IQGPU_AGC_BLOCK_TC=$tc timeout 300 python bench.py --workload cfg4 --steps 10 $B > gpurun_out/r2c_bench_cfg4_tc$tc.json 2> gpurun_out/r2c_bench_cfg4_tc$tc.err
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:agc_rms -s 2 -c 1 -o gpurun_out/r2c_agcrms_cfg4 python bench.py --workload cfg4 $N > gpurun_out/r2c_ncu_cfg4.log 2>&1
ls -la gpurun_out | tail -8
