#!/bin/bash
# Round 2, call C: what an LDG.256 costs in L1 wavefronts (microbenchmark under ncu), the warp-per-ray experiment,
# cache-hint variants of the record load.
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 120 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__lsu_writeback_active.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_ldg256_wavefronts.csv tools/ldg256_wavefronts > gpurun_out/${tag}_ldg256.log 2>&1
tail -2 gpurun_out/${tag}_ldg256.log
timeout 200 python tools/ab_quick.py --steps 4 --flags "0,32768,32768:fast_coop=1" > gpurun_out/${tag}_ab_coop.json 2> gpurun_out/${tag}_ab_coop.err
python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_ab_coop.json'))
for k,v in d.items():
    if k.startswith('flags_'): print(k, v['ms_min'], v['frames_per_s'], v['max_rel_diff_vs_first'])
"
tail -2 gpurun_out/${tag}_ab_coop.err
for m in plain noalloc evlast evfirst; do
BIRP_B200_LIB=$PWD/tools/scratch/libbirp_b200_$m.so timeout 120 python tools/ab_quick.py --steps 4 --flags "0,32768" > gpurun_out/${tag}_ab_load_$m.json 2> gpurun_out/${tag}_ab_load_$m.err
python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_ab_load_$m.json'))
for k,v in d.items():
    if k.startswith('flags_'): print('$m', k, v['ms_min'], v['frames_per_s'], v['max_rel_diff_vs_first'])
"
done
