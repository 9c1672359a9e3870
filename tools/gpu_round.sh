#!/bin/bash
# One GPU round-trip: parity tests, smoke, bench, ncu launch list + full captures.
# usage (under gpurun): [NCU=0|1] [TESTS=0|1] bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/smi_$TAG.txt
if [ "${TESTS:-1}" = "1" ]; then
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) | tee gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$TAG.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks","roofline")})
print(d["cpu_baseline"])
for r in d["kernels"]: print({k:v for k,v in r.items() if k!="note"})
PY
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
   --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py --steps 4 > gpurun_out/ncu_list_$TAG.log 2>&1
# index_max is a standalone API op (not in the classifier step): bench.py's standalone section launches it
timeout 600 ncu --set full --clock-control none --import-source on -k regex:index_max -s 1 -c 1 \
   -o gpurun_out/prof_index_max_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_im_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pointresnet_tc -s 2 -c 1 \
   -o gpurun_out/prof_pointresnet_tc_$TAG -f python tools/profile_step.py --steps 3 > gpurun_out/ncu_tc_$TAG.log 2>&1
tail -3 gpurun_out/ncu_tc_$TAG.log
ls -la gpurun_out/ | grep $TAG
fi
