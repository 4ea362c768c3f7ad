set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2l_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for w in cfg2 cfg1 cfg5 cfg3 cfg4; do
  IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2l_bench_$w.json 2> gpurun_out/r2l_bench_$w.err
done
for m in 0 1 2; do
  IQGPU_ARB_PAIRS=$m timeout 300 python bench.py --workload cfg2 --steps 20 $B > gpurun_out/r2l_bench_cfg2_pairs$m.json 2> gpurun_out/r2l_bench_cfg2_pairs$m.err
  IQGPU_ARB_PAIRS=$m timeout 300 python bench.py --workload cfg5 --steps 20 $B > gpurun_out/r2l_bench_cfg5_pairs$m.json 2> gpurun_out/r2l_bench_cfg5_pairs$m.err
done
IQGPU_NO_DC_FOLD=1 timeout 300 python bench.py --workload cfg2 --steps 20 $B > gpurun_out/r2l_bench_cfg2_nofold.json 2> gpurun_out/r2l_bench_cfg2_nofold.err
N="--steps 2 --warmup 1 $B"
for m in 0 2; do
IQGPU_ARB_PAIRS=$m timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 1 -c 1 -o gpurun_out/r2l_ff2_cfg2_m$m python bench.py --workload cfg2 $N > gpurun_out/r2l_ncu_cfg2_m$m.log 2>&1
done
ls -la gpurun_out | tail -4
