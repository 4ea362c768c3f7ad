set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2f_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
for w in cfg1 cfg3 cfg4 cfg5 k1; do
  IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 4 -c 1 -o gpurun_out/r2f_ff2_cfg2 python bench.py --workload cfg2 $N > gpurun_out/r2f_ncu_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 5 -c 1 -o gpurun_out/r2f_ff2_cfg1 python bench.py --workload cfg1 $N > gpurun_out/r2f_ncu_cfg1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_cfg2.csv python bench.py --workload cfg2 $N > gpurun_out/r2f_l_cfg2.log 2>&1
ls -la gpurun_out | tail -6
