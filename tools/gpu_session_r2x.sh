set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 3 --sharded-capture= --stage-leg="
for w in cfg1 cfg3 cfg4 cfg5 k1; do
  timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2x_bench_$w.json 2> gpurun_out/r2x_bench_$w.err
done
timeout 900 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err
timeout 600 python bench.py --workload file:cfg2 --samples 1073741824 --steps 3 --warmup 1 > gpurun_out/r2x_file_cfg2.json 2> gpurun_out/r2x_file_cfg2.err
K='regex:^(agc_|arb_|dc_|fft|fir_|fused_|halfband|iq_opt|post_|pre_|w2_)'
Q="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
for w in cfg3 cfg4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2x_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 1 $Q > gpurun_out/r2x_l_$w.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 6 -c 1 -o gpurun_out/r2x_ff2_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 1 $Q > gpurun_out/r2x_ncu_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 6 -c 1 -o gpurun_out/r2x_ff2_cfg1 python bench.py --workload cfg1 --steps 2 --warmup 1 $Q > gpurun_out/r2x_ncu_cfg1.log 2>&1
du -sh gpurun_out
