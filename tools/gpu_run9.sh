#!/bin/bash
# every stage has a tight timeout; a canary aborts the whole script if the megakernel hangs or faults
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
run build 300 python __graft_entry__.py || exit 1
run canary 150 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider -x -k "megakernel" || { echo "CANARY FAILED - aborting"; exit 1; }
run canary2 240 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider -x -k "chat_flow or batch_invariance" || { echo "CANARY2 FAILED - aborting"; exit 1; }
run pytest_gpu 420 python -m pytest tests -m gpu -q -p no:cacheprovider
TAILN=22 run trace_nol2 420 env MINIGPT4_B200_L2_AHEAD=0 python tools/mega_trace.py || exit 1
TAILN=22 run trace 300 python tools/mega_trace.py || exit 1
TAILN=30 run bench 600 python bench.py --steps 2 --warmup 3 --no-cpu
