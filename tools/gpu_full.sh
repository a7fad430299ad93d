#!/bin/bash
# Full 1-GPU check-out: tests, bench, prefill profile.   gpurun --timeout 2400 -- 'bash tools/gpu_full.sh'
mkdir -p gpurun_out
TAG=${TAG:-r2_v6}
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary6 200 python tools/canary.py || exit 1
NOTRACE=1 MINIGPT4_B200_MEGA_GEN=4 TAILN=1 run ab_v4 120 python tools/mega_trace.py
NOTRACE=1 TAILN=1 run ab_v6 120 python tools/mega_trace.py
TAILN=3 run prefill_plain 300 python tools/prof_prefill.py
TAILN=8 run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider
REPS=1 TAILN=2 run ncu_prefill_full 600 ncu --set full --clock-control none --import-source on -k regex:prefill_gemm -s 6 -c 2 -o gpurun_out/${TAG}_prefill -f python tools/prof_prefill.py
TAILN=3 run bench 600 python bench.py
echo done
