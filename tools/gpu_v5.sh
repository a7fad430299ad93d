#!/bin/bash
# Generation-5 megakernel: correctness ladder (canary, bounded spins + timeouts), then timing A/B and per-op trace on the 4-layer 7B-wide model.
#   gpurun --timeout 900 -- 'bash tools/gpu_v5.sh'
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
export MINIGPT4_B200_MEGA5=1
TAILN=5 run canary5 200 python tools/canary.py
MINIGPT4_B200_MEGA5_NOREG=1 MINIGPT4_B200_L2_AHEAD=32 TAILN=5 run canary5_noreg_l2 200 python tools/canary.py
MINIGPT4_B200_MEGA5= NOTRACE=1 TAILN=1 run ab_v4 120 python tools/mega_trace.py
NOTRACE=1 TAILN=1 run ab5 120 python tools/mega_trace.py
for a in 16 48 128; do MINIGPT4_B200_L2_AHEAD=$a NOTRACE=1 TAILN=1 run ab5_l2_$a 120 python tools/mega_trace.py; done
TAILN=24 run trace5 200 python tools/mega_trace.py
MINIGPT4_B200_L2_AHEAD=48 TAILN=24 run trace5_l2 200 python tools/mega_trace.py
echo done
