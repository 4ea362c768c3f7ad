set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2p_smoke.log 2>&1
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
for w in cfg1 cfg2 cfg3 cfg4 cfg5 k1; do
  IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2p_bench_$w.json 2> gpurun_out/r2p_bench_$w.err
done
timeout 900 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err
timeout 600 python bench.py --workload file:cfg2 --samples 1073741824 --steps 3 --warmup 1 > gpurun_out/r2p_file_cfg2.json 2> gpurun_out/r2p_file_cfg2.err
K='regex:^(agc_|arb_|dc_|fft|fir_|fused_|halfband|iq_opt|post_|pre_|w2_)'
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2p_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 1 $B > gpurun_out/r2p_l_$w.log 2>&1
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fftfilt2 -s 2 -c 1 -o gpurun_out/r2p_fft_cfg3 python bench.py --workload cfg3 $N > gpurun_out/r2p_ncu_cfg3.log 2>&1
du -sh gpurun_out
