#!/bin/bash
# Generation-5 megakernel with group slots.   gpurun --timeout 1200 -- 'bash tools/gpu_v5c.sh'
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
export MINIGPT4_B200_MEGA5=1
TAILN=4 run canary5 200 python tools/canary.py
MINIGPT4_B200_MEGA_FLAGS=5 MINIGPT4_B200_MEGA5_NOREG=1 TAILN=4 run canary5_f5_noreg 200 python tools/canary.py
export NOTRACE=1 TAILN=1
MINIGPT4_B200_MEGA5= run ab_v4 120 python tools/mega_trace.py
for f in 1 5; do MINIGPT4_B200_MEGA_FLAGS=$f run ab5_f$f 120 python tools/mega_trace.py; done
for i in 2 3 5 6; do MINIGPT4_B200_INFLIGHT=$i run ab5_if$i 120 python tools/mega_trace.py; done
export TRACE_LAYERS=1 TRACE_VOCAB=200000
MINIGPT4_B200_MEGA5= run stream_v4 300 python tools/mega_trace.py
run stream5 120 python tools/mega_trace.py
MINIGPT4_B200_INFLIGHT=3 run stream5_if3 120 python tools/mega_trace.py
MINIGPT4_B200_INFLIGHT=5 run stream5_if5 120 python tools/mega_trace.py
unset TRACE_LAYERS TRACE_VOCAB NOTRACE
TAILN=20 run trace5 200 python tools/mega_trace.py
unset MINIGPT4_B200_MEGA5
TAILN=15 run pytest_prefill 300 python -m pytest tests/test_prefill_gpu.py -m gpu -x -q -p no:cacheprovider
echo done
