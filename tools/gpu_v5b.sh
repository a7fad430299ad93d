#!/bin/bash
# Generation-5 megakernel, variant matrix.   gpurun --timeout 1200 -- 'bash tools/gpu_v5b.sh'
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
export MINIGPT4_B200_MEGA5=1
TAILN=4 run canary5 200 python tools/canary.py
MINIGPT4_B200_MEGA_FLAGS=13 TAILN=4 run canary5_f13 200 python tools/canary.py
export NOTRACE=1 TAILN=1
MINIGPT4_B200_MEGA5= run ab_v4 120 python tools/mega_trace.py
for f in 1 5 9 13; do MINIGPT4_B200_MEGA_FLAGS=$f run ab5_f$f 120 python tools/mega_trace.py; done
MINIGPT4_B200_MEGA_FLAGS=9 MINIGPT4_B200_L2_AHEAD=48 run ab5_f9_l2 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=9 MINIGPT4_B200_INFLIGHT=13 run ab5_f9_if13 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=1 MINIGPT4_B200_INFLIGHT=13 run ab5_f1_if13 120 python tools/mega_trace.py
# pure streaming: 1 layer + a 200k-row output matrix (819 MB of 852 MB per token in ONE barrier-free op)
export TRACE_LAYERS=1 TRACE_VOCAB=200000
MINIGPT4_B200_MEGA5= run stream_v4 300 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=1 run stream5_f1 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=9 run stream5_f9 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=9 MINIGPT4_B200_INFLIGHT=13 run stream5_f9_if13 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=9 MINIGPT4_B200_L2_AHEAD=48 run stream5_f9_l2 120 python tools/mega_trace.py
unset TRACE_LAYERS TRACE_VOCAB NOTRACE
MINIGPT4_B200_MEGA_FLAGS=9 TAILN=20 run trace5_f9 200 python tools/mega_trace.py
unset MINIGPT4_B200_MEGA5
TAILN=15 run pytest_prefill 300 python -m pytest tests/test_prefill_gpu.py -m gpu -x -q -p no:cacheprovider
echo done
