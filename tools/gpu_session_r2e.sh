set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./tools/ubench/_build/dlat > gpurun_out/r2e_dlat.txt 2>&1
IQGPU_ARB_PAIRS=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "fused or cfg1 or cfg5 or golden or chunk_train or cfg3" 2>&1 | tail -30 > gpurun_out/r2e_pytest_quad.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture="
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do
  IQGPU_VERBOSE=1 timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2e_bench_$w.json 2> gpurun_out/r2e_bench_$w.err
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:agc_rms -s 2 -c 1 -o gpurun_out/r2e_agcrms_cfg4 python bench.py --workload cfg4 $N > gpurun_out/r2e_ncu_cfg4.log 2>&1
IQGPU_ARB_PAIRS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 3 -c 1 -o gpurun_out/r2e_ff2_cfg1 python bench.py --workload cfg1 $N > gpurun_out/r2e_ncu_cfg1.log 2>&1
ls -la gpurun_out | tail -6
