set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
Q="--no-cpu-baseline --sustained-seconds 0 --no-pcie-probe --e2e-steps 0 --sharded-capture= --stage-leg="
IQGPU_ARB_PAIRS=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_front2 -s 1 -c 1 -o gpurun_out/r3e_ff2_cfg3 python bench.py --workload cfg3 --steps 2 --warmup 1 $Q > gpurun_out/r3e_ncu_cfg3.log 2>&1
