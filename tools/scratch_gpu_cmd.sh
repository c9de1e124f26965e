# GPU check 3: right-sized dynamic ray table (more L1) vs the static 8 KB table, tile shapes under oct records
mkdir -p gpurun_out
timeout 100 python tools/ab_time.py --workload orbit --flags 0,8192 --steps 6 > gpurun_out/r1z3_ab_dyn.json 2> gpurun_out/r1z3_ab_dyn.err
cat gpurun_out/r1z3_ab_dyn.json
BIRP_B200_LIB=$PWD/tools/scratch/libbirp_b200_static.so timeout 100 python tools/ab_time.py --workload orbit --flags 0,8192 --steps 6 > gpurun_out/r1z3_ab_static.json 2> gpurun_out/r1z3_ab_static.err
cat gpurun_out/r1z3_ab_static.json
for t in 2 4; do
BIRP_TILE_W_LOG2=$t timeout 100 python tools/ab_time.py --workload orbit --flags 0 --steps 6 > gpurun_out/r1z3_ab_tile$t.json 2> gpurun_out/r1z3_ab_tile$t.err
cat gpurun_out/r1z3_ab_tile$t.json
done
timeout 100 python tools/ab_time.py --workload nearest --flags 0 --steps 6 > gpurun_out/r1z3_ab_nearest.json 2> gpurun_out/r1z3_ab_nearest.err
cat gpurun_out/r1z3_ab_nearest.json
timeout 100 python tools/ab_time.py --workload hires256 --flags 0,8192 --steps 4 > gpurun_out/r1z3_ab_hires256.json 2> gpurun_out/r1z3_ab_hires256.err
cat gpurun_out/r1z3_ab_hires256.json
