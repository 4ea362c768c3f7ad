set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
for tc in 40 30 20 15 10 7; do
  IQGPU_DEBUG_AGC=1 IQGPU_AGC_BLOCK_TC=$tc timeout 300 python bench.py --workload cfg4 --steps 10 $B > gpurun_out/r2v_bench_cfg4_tc$tc.json 2> gpurun_out/r2v_bench_cfg4_tc$tc.err
done
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -k "agc" 2>&1 | tail -5 > gpurun_out/r2v_pytest.log
