set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -80 > gpurun_out/r2b_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for w in cfg2 cfg1 cfg3 cfg5; do
  timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err
  IQGPU_NO_RAW_TMA=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2b_bench_${w}_notma.json 2> gpurun_out/r2b_bench_${w}_notma.err
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 5 -c 1 -o gpurun_out/r2b_ff2_cfg2 python bench.py --workload cfg2 $N > gpurun_out/r2b_ncu_cfg2.log 2>&1
ls -la gpurun_out | tail -12
