set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/r2z_pytest.log
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
for w in cfg1 cfg3 cfg4; do
  timeout 300 python bench.py --workload $w --steps 20 $B > gpurun_out/r2z_bench_$w.json 2> gpurun_out/r2z_bench_$w.err
done
timeout 900 python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err
# memory checker over the kernels that changed this round (small inputs; the tool slows kernels 10-50x)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x --timeout 1400 -k "every_cascade_depth or fft_filter_equals_fir or rms_agc_parity or polyphase_variants or fir_filter_stage or golden_fixture or all_input_formats" > gpurun_out/r2z_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2z_memcheck.log
tail -c 3000 gpurun_out/r2z_memcheck.log > gpurun_out/r2z_memcheck_tail.log
du -sh gpurun_out
