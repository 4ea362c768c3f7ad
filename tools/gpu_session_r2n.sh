set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2n_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
for w in cfg5 cfg1; do
  timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2n_bench_$w.json 2> gpurun_out/r2n_bench_$w.err
done
for tc in 256 512 1024; do timeout 300 python bench.py --workload file:cfg2 --samples 536870912 --steps 3 --warmup 1 --no-cpu-baseline --train-chunks $tc > gpurun_out/r2n_file_cfg2_tc$tc.json 2> gpurun_out/r2n_file_cfg2_tc$tc.err; done
timeout 900 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err
K='regex:^(agc_|arb_|dc_|fft|fir_|fused_|halfband|iq_opt|post_|pre_|w2_)'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2n_launches_cfg5.csv python bench.py --workload cfg5 --steps 3 --warmup 1 $B > gpurun_out/r2n_l_cfg5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 4 -c 1 -o gpurun_out/r2n_ff2_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 1 $B > gpurun_out/r2n_ncu_cfg5.log 2>&1
ls -la gpurun_out | tail -4
