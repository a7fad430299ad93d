#!/bin/bash
# Generation-6 megakernel check-out.   gpurun --timeout 1500 -- 'bash tools/gpu_v6.sh'
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary6 200 python tools/canary.py || exit 1
MINIGPT4_B200_MEGA_FLAGS=3 TAILN=4 run canary6_f3 200 python tools/canary.py
export NOTRACE=1 TAILN=1
run ab_v6_f1 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=3 run ab_v6_f3 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=3 MINIGPT4_B200_L2_WINDOW=393216 run ab_v6_f3_w384k 120 python tools/mega_trace.py
MINIGPT4_B200_MEGA_FLAGS=3 MINIGPT4_B200_L2_WINDOW=98304 run ab_v6_f3_w96k 120 python tools/mega_trace.py
unset NOTRACE
TAILN=6 run pytest_gpu 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider
TAILN=3 run bench 500 python bench.py
TAILN=3 run bench_q5k 700 python bench.py --wtype q5_k --tokens 256 --steps 2 --warmup 1
echo done
