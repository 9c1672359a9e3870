#!/bin/bash
# One GPU round-trip (round 2). Sections are selected with env vars (1 = run):
#   TESTS   pytest -m gpu + smoke          BENCH   bench.py default run (+ reference arm REFARM=1)
#   LONG    bench.py --steps $LONG (sustained-rate evidence, clocks/power trace)
#   LIST    ncu launch list of 4 eager steps
#   NCU     space-separated kernel regexes for `ncu --set full` captures of tools/profile_step.py
#   NCUB    same, but captured from bench.py's standalone section (index_max, som_assign+mask)
#   SANI    compute-sanitizer memcheck + racecheck + synccheck of a B=2, N=5000 forward
#   CONFIGS tools/bench_configs.py            EXTRA   any shell command
# usage (under gpurun): TESTS=1 BENCH=1 bash tools/gpu_round2.sh r02a
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $O/smi_$TAG.txt
if [ "${TESTS:-0}" = "1" ]; then
  (timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} 2>&1 | tail -40) | tee $O/pytest_gpu_$TAG.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke_$TAG.log
fi
if [ "${BENCH:-0}" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
  tail -3 $O/bench_$TAG.err
  python - <<PY
import json
d=json.load(open("$O/bench_$TAG.json"))
print({k:d.get(k) for k in ("value","ms_per_step","e2e","gpu_launches","clocks","parity")})
print({k:v for k,v in d["roofline"].items() if k!="secondary"})
for r in d["roofline"].get("secondary",[]): print(r)
print(d["cpu_baseline"])
for r in d["kernels"]: print({k:v for k,v in r.items() if k!="note"})
PY
fi
if [ "${REFARM:-0}" = "1" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err
  cat $O/bench_ref_$TAG.json | cut -c1-600
fi
if [ -n "${LONG:-}" ]; then
  timeout 900 python bench.py --steps $LONG --warmup 20 --no-cpu-baseline > $O/bench_long_$TAG.json 2> $O/bench_long_$TAG.err
  python - <<PY
import json
d=json.load(open("$O/bench_long_$TAG.json"))
print("LONG", {k:d.get(k) for k in ("value","ms_per_step","steps","e2e","clocks","step_ms")})
PY
fi
if [ "${LIST:-0}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
     --log-file $O/launches_$TAG.csv python tools/profile_step.py --steps 4 > $O/ncu_list_$TAG.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
     --log-file $O/launches_cfg4_$TAG.csv python tools/profile_step.py --steps 3 --task autoencoder --batch 32 > $O/ncu_list_cfg4_$TAG.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
     --log-file $O/launches_cfg3_$TAG.csv python tools/profile_step.py --steps 3 --task segmenter --batch 32 --npts 1024 > $O/ncu_list_cfg3_$TAG.log 2>&1
fi
for K in ${NCU:-}; do
  N=$(echo $K | tr -c 'a-zA-Z0-9_\n' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-2} -c 1 \
     -o $O/prof_${N}_$TAG -f python tools/profile_step.py --steps 3 ${PROFILE_ARGS:-} > $O/ncu_${N}_$TAG.log 2>&1
  tail -2 $O/ncu_${N}_$TAG.log
done
for K in ${NCUB:-}; do
  N=$(echo $K | tr -c 'a-zA-Z0-9_\n' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 \
     -o $O/prof_${N}_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_${N}_$TAG.log 2>&1
  tail -2 $O/ncu_${N}_$TAG.log
done
# NCUO: "op:kernel_regex" pairs captured from tools/profile_ops.py (standalone ops)
for PAIR in ${NCUO:-}; do
  OP=${PAIR%%:*}; K=${PAIR##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCUO_SKIP:-2} -c 1 \
     -o $O/prof_${OP}_$TAG -f python tools/profile_ops.py --op $OP > $O/ncu_${OP}_$TAG.log 2>&1
  tail -2 $O/ncu_${OP}_$TAG.log
done
# OPS: standalone ops timed with CUDA events (tools/profile_ops.py)
for OP in ${OPS:-}; do python tools/profile_ops.py --op $OP --reps 6 2>&1 | tail -1; done
if [ "${SANI:-0}" = "1" ]; then
  for TOOL in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_all.py \
       > $O/sanitizer_${TOOL}_$TAG.log 2>&1
    echo "== $TOOL"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" $O/sanitizer_${TOOL}_$TAG.log | tail -3
  done
fi
if [ "${CONFIGS:-0}" = "1" ]; then
  timeout 900 python tools/bench_configs.py > $O/bench_configs_$TAG.json 2> $O/bench_configs_$TAG.err
  tail -5 $O/bench_configs_$TAG.err; cut -c1-1500 $O/bench_configs_$TAG.json
fi
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA"; fi
ls -la $O/ | grep $TAG
