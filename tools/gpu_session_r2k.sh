set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2k_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for w in cfg2 cfg1 cfg5 cfg3 cfg4; do
  IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2k_bench_$w.json 2> gpurun_out/r2k_bench_$w.err
done
for w in cfg2 cfg5; do
  IQGPU_ARB_PAIRS=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2k_bench_${w}_pairs1.json 2> gpurun_out/r2k_bench_${w}_pairs1.err
  IQGPU_ARB_PAIRS=2 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2k_bench_${w}_pairs2.json 2> gpurun_out/r2k_bench_${w}_pairs2.err
done
N="--steps 2 --warmup 1 $B"
IQGPU_ARB_PAIRS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 1 -c 1 -o gpurun_out/r2k_ff2_cfg2 python bench.py --workload cfg2 $N > gpurun_out/r2k_ncu_cfg2.log 2>&1
IQGPU_ARB_PAIRS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 1 -c 1 -o gpurun_out/r2k_ff2_cfg1 python bench.py --workload cfg1 $N > gpurun_out/r2k_ncu_cfg1.log 2>&1
K='regex:^(agc_|arb_|dc_|fft|fir_|fused_|halfband|iq_opt|post_|pre_|w2_)'
for w in cfg2 cfg4 cfg5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2k_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 1 $B > gpurun_out/r2k_l_$w.log 2>&1
done
timeout 600 python bench.py --workload file:cfg2 --samples 536870912 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_file_cfg2.json 2> gpurun_out/r2k_file_cfg2.err
ls -la gpurun_out | tail -4
