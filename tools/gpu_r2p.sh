#!/bin/bash
# Round 2, call P (1 GPU): CMEM kernel -- warps per CTA (8 / 4 / 2) x register cap (64 / 48 / 40), default and tolerance mode.
tag=${1:-r2p}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k 'cmem and not never_disagrees' > gpurun_out/${tag}_pytest_cmem.log 2>&1
tail -2 gpurun_out/${tag}_pytest_cmem.log
for v in default cm3 cm5 cm6; do
  lib=birp_cpp_b200/libbirp_b200.so; [ $v != default ] && lib=tools/scratch/libbirp_b200_$v.so
  for w in 8 4 2; do
    BIRP_B200_LIB=$PWD/$lib BIRP_CTA_WARPS=$w timeout 120 python tools/ab_time.py --workload cmem --flags 0,4 --steps 3 > gpurun_out/${tag}_ab_cmem_${v}_w$w.json 2> gpurun_out/${tag}_ab_cmem_${v}_w$w.err
    python -c "
import json
d=json.load(open('gpurun_out/${tag}_ab_cmem_${v}_w$w.json'))
print('$v', 'w$w', d['flags_0']['ms_min'], d['flags_4']['ms_min'], d['flags_4']['pixels_within_1e-5_of_first'])
" || tail -2 gpurun_out/${tag}_ab_cmem_${v}_w$w.err
  done
done
