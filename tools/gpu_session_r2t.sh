set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2t_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
for w in cfg3 cfg4; do
  timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2t_bench_$w.json 2> gpurun_out/r2t_bench_$w.err
done
N="--steps 2 --warmup 1 $B"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fftfilt2 -s 2 -c 1 -o gpurun_out/r2t_fft_cfg3 python bench.py --workload cfg3 $N > gpurun_out/r2t_ncu_cfg3.log 2>&1
IQGPU_FFT_NO_PREFETCH=1 timeout 300 python bench.py --workload cfg3 --steps 20 $B > gpurun_out/r2t_bench_cfg3_nopf.json 2> gpurun_out/r2t_bench_cfg3_nopf.err
du -sh gpurun_out
