set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x --timeout 1400 -k "every_cascade_depth or fft_filter_equals_fir or fir_filter_stage or polyphase_variants_are_bit_identical_at_every_rate_class" > gpurun_out/r3d_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r3d_racecheck.log
tail -c 4000 gpurun_out/r3d_racecheck.log > gpurun_out/r3d_racecheck_tail.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x --timeout 800 -k "every_cascade_depth or rms_agc_parity or digital_agc" > gpurun_out/r3d_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/r3d_synccheck.log
tail -c 3000 gpurun_out/r3d_synccheck.log > gpurun_out/r3d_synccheck_tail.log
rm -f gpurun_out/r3d_racecheck.log gpurun_out/r3d_synccheck.log
