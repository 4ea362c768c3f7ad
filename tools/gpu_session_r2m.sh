set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2m_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for w in cfg2 cfg5 cfg1; do
  IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2m_bench_$w.json 2> gpurun_out/r2m_bench_$w.err
done
IQGPU_NO_DC_FOLD=1 timeout 300 python bench.py --workload cfg2 --steps 20 $B > gpurun_out/r2m_bench_cfg2_nofold.json 2> gpurun_out/r2m_bench_cfg2_nofold.err
K='regex:^(agc_|arb_|dc_|fft|fir_|fused_|halfband|iq_opt|post_|pre_|w2_)'
for w in cfg2 cfg5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2m_launches_$w.csv python bench.py --workload $w --steps 3 --warmup 1 $B > gpurun_out/r2m_l_$w.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fir_param -s 1 -c 1 -o gpurun_out/r2m_fir_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 1 $B > gpurun_out/r2m_ncu_fir.log 2>&1
ls -la gpurun_out | tail -4
