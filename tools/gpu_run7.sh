#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run build 600 python __graft_entry__.py
run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider
TAILN=30 run trace 900 python tools/mega_trace.py
TAILN=30 run bench 1500 python bench.py --steps 2 --warmup 3
