set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
E="--no-cpu-baseline --sustained-seconds 0 --sharded-capture= --stage-leg= --steps 5 --e2e-steps 3"
timeout 300 python bench.py $E > gpurun_out/r2o_e2e_new.json 2> gpurun_out/r2o_e2e_new.err
(cd _old && timeout 300 python bench.py --no-cpu-baseline --sustained-seconds 0 --sharded-capture= --steps 5 --e2e-steps 3 > ../gpurun_out/r2o_e2e_old.json 2> ../gpurun_out/r2o_e2e_old.err)
IQGPU_NO_DC_FOLD=1 timeout 300 python bench.py $E > gpurun_out/r2o_e2e_nofold.json 2> gpurun_out/r2o_e2e_nofold.err
IQGPU_ARB_PAIRS=0 timeout 300 python bench.py $E > gpurun_out/r2o_e2e_pairs0.json 2> gpurun_out/r2o_e2e_pairs0.err
timeout 300 python bench.py $E > gpurun_out/r2o_e2e_new2.json 2> gpurun_out/r2o_e2e_new2.err
for tc in 64 128 256; do timeout 300 python bench.py --workload file:cfg2 --samples 536870912 --steps 3 --warmup 1 --no-cpu-baseline --train-chunks $tc > gpurun_out/r2o_file_cfg2_tc$tc.json 2> gpurun_out/r2o_file_cfg2_tc$tc.err; done
ls -la gpurun_out | tail -3
