#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
MINIGPT4_B200_MEGA_FLAGS=3 TAILN=4 run canary 240 python tools/canary.py || { echo "CANARY FAILED - aborting"; exit 1; }
NOTRACE=1 MINIGPT4_B200_MEGA_FLAGS=1 TAILN=1 run ab_f1 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_MEGA_FLAGS=3 TAILN=1 run ab_f3 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_MEGA_FLAGS=2 TAILN=1 run ab_f2 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_MEGA_FLAGS=0 TAILN=1 run ab_f0 120 python tools/mega_trace.py
M=dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
NOTRACE=1 MINIGPT4_B200_MEGA_FLAGS=3 TAILN=8 run ncu_f3 200 ncu --metrics $M --clock-control none -k regex:decode_megakernel -s 20 -c 1 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=3 TAILN=22 run trace_f3 200 python tools/mega_trace.py
echo done
