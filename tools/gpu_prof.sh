#!/bin/bash
# Lean GPU-box run: megakernel canary, per-op trace, A/B of the tuning knobs, bench, ncu --set full captures.
mkdir -p gpurun_out
TAG=${TAG:-r1_v3}
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary 240 python tools/canary.py || { echo "CANARY FAILED - aborting"; exit 1; }
TAILN=22 run trace 200 python tools/mega_trace.py
NOTRACE=1 TAILN=1 run ab_default 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_MEGA_FLAGS=0 TAILN=1 run ab_nokvprefetch 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_INFLIGHT2=10 TAILN=1 run ab_inflight2_10 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_L2_AHEAD=24 TAILN=1 run ab_l2ahead24 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_L2_AHEAD=96 TAILN=1 run ab_l2ahead96 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_INFLIGHT=6 TAILN=1 run ab_inflight6 120 python tools/mega_trace.py
TAILN=2 run ncu_mega 420 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 10 -c 2 -o gpurun_out/${TAG}_mega -f python tools/prof_decode.py
TAILN=3 run bench 600 python bench.py --steps 3 --warmup 3 --no-cpu
TAILN=2 run ncu_vision 420 ncu --set full --clock-control none --import-source on -k 'regex:gemm_f16|attention_f32|layernorm' -s 465 -c 12 -o gpurun_out/${TAG}_vision -f python tools/prof_vision.py
echo done
