#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
run build 300 python __graft_entry__.py || exit 1
run canary 280 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider -x -k "megakernel" || { echo "CANARY FAILED - aborting"; exit 1; }
TAILN=22 run trace 300 python tools/mega_trace.py || exit 1
NOTRACE=1 TAILN=2 run notrace 300 python tools/mega_trace.py
