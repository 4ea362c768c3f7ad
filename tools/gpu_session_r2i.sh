set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2i_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for w in cfg2 cfg1 cfg3 cfg4 cfg5; do
  IQGPU_DEBUG_AGC=1 IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2i_bench_$w.json 2> gpurun_out/r2i_bench_$w.err
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 4 -c 1 -o gpurun_out/r2i_ff2_cfg2 python bench.py --workload cfg2 $N > gpurun_out/r2i_ncu_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:agc_rms -s 2 -c 1 -o gpurun_out/r2i_agcrms_cfg4 python bench.py --workload cfg4 $N > gpurun_out/r2i_ncu_cfg4.log 2>&1
ls -la gpurun_out | tail -4
