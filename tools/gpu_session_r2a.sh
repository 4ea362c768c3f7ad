set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
for w in cfg1 cfg3 cfg4 cfg5 k1; do
  timeout 300 python bench.py --workload $w --steps 20 --sharded-capture '' $B > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err
done
N="--steps 2 --warmup 1 --e2e-steps 0 --sharded-capture= $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 5 -c 1 -o gpurun_out/r2a_ff2_cfg1 python bench.py --workload cfg1 $N > gpurun_out/r2a_ncu_cfg1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:agc_rms -s 2 -c 1 -o gpurun_out/r2a_agcrms_cfg4 python bench.py --workload cfg4 $N > gpurun_out/r2a_ncu_cfg4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pre_kernel -s 2 -c 1 -o gpurun_out/r2a_pre_k1 python bench.py --workload k1 $N > gpurun_out/r2a_ncu_k1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches_cfg4.csv python bench.py --workload cfg4 $N > gpurun_out/r2a_l_cfg4.log 2>&1
ls -la gpurun_out | tail -30
