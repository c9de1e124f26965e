#!/bin/bash
# Round 2, call O (N GPUs): closing multi-GPU lines of the committed binary (small first and last chunks in every many-frame batch, 4-warp CTAs for large images).
tag=${1:-r2j}
N=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest_${N}gpu.log 2>&1
tail -3 gpurun_out/${tag}_pytest_${N}gpu.log
for spec in "orbit:0" "hires512x4k:0" "orbit:32768"; do
w=${spec%%:*}; fl=${spec#*:}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --extra-flags $fl > gpurun_out/${tag}_bench_${w}_flags${fl}_${N}gpu.json 2> gpurun_out/${tag}_bench_${w}_flags${fl}_${N}gpu.err; echo "bench $w $fl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_${w}_flags${fl}_${N}gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['bit_exact'], d['gather_check'], d['e2e_gather_check'], d['failed'], d.get('render_only_ms_per_rank'), (d.get('tolerance_mode') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
grep -v "warnings.warn\|resource_tracker\|^W1\|OMP_NUM\|^\*\*\*" gpurun_out/${tag}_bench_${w}_flags${fl}_${N}gpu.err | tail -3
done
