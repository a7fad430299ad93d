#!/bin/bash
# GPU-box run: canary, L2 look-ahead A/B (time + DRAM bytes), new quantised-container tests, bench under both settings.
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary 240 python tools/canary.py || { echo "CANARY FAILED - aborting"; exit 1; }
NOTRACE=1 TAILN=1 run ab_l2_48 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_L2_AHEAD=0 TAILN=1 run ab_l2_0 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_L2_AHEAD=16 TAILN=1 run ab_l2_16 120 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_L2_AHEAD=128 TAILN=1 run ab_l2_128 120 python tools/mega_trace.py
M=dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
NOTRACE=1 TAILN=8 run ncu_l2_48 200 ncu --metrics $M --clock-control none -k regex:decode_megakernel -s 20 -c 1 python tools/mega_trace.py
NOTRACE=1 MINIGPT4_B200_L2_AHEAD=0 TAILN=8 run ncu_l2_0 200 ncu --metrics $M --clock-control none -k regex:decode_megakernel -s 20 -c 1 python tools/mega_trace.py
MINIGPT4_B200_L2_AHEAD=0 TAILN=22 run trace_l2_0 200 python tools/mega_trace.py
TAILN=6 run pytest_qvision 400 python -m pytest tests/test_quantized_vision.py -m gpu -x -q -p no:cacheprovider
TAILN=2 run bench_l2_48 500 python bench.py --steps 3 --warmup 3 --no-cpu
MINIGPT4_B200_L2_AHEAD=0 TAILN=2 run bench_l2_0 500 python bench.py --steps 3 --warmup 3 --no-cpu
echo done
