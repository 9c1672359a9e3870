#!/bin/bash
# Multi-GPU round-trip (run under `gpurun --gpus N`): NCCL parity test (N>=2) + bench at N ranks with
# both collectives. usage: bash tools/multi_gpu_round.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv | head -10
python -m pytest tests/test_gpu_nccl2.py -m gpu -q 2>&1 | tail -3 | tee $O/pytest_nccl2_$TAG.log
for COLL in ${COLLS:-torch sonet}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 20 --warmup 3 --collective $COLL > $O/bench_${N}gpu_${COLL}_$TAG.json 2> $O/bench_${N}gpu_${COLL}_$TAG.err
  tail -2 $O/bench_${N}gpu_${COLL}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_${N}gpu_${COLL}_$TAG.json") if l.startswith("{")][-1])
    print("$COLL", {k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","gather_check")})
except Exception as e: print("no bench line", e)
PY
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_$TAG.json 2> $O/bench_1gpu_$TAG.err
python -c "
import json;d=json.load(open('$O/bench_1gpu_$TAG.json'));print('1gpu',d['value'],d['ms_per_step'])"
