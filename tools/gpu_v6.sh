#!/bin/bash
# Generation-6 megakernel check-out.   gpurun --timeout 1200 -- 'bash tools/gpu_v6.sh'
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary6 200 python tools/canary.py || exit 1
export NOTRACE=1 TAILN=1
MINIGPT4_B200_L2_WINDOW=0 run ab_v6_w0 120 python tools/mega_trace.py
run ab_v6_w192k 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=0 run ab_v6_f0 120 python tools/mega_trace.py
unset NOTRACE
TAILN=30 run trace6 200 python tools/mega_trace.py
TAILN=3 run bench 500 python bench.py
echo done
