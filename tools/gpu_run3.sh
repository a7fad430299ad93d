#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run build 600 python __graft_entry__.py
run mega_test 300 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider -k "megakernel or batch_invariance"
run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x
TAILN=30 run bench 1500 python bench.py --steps 2 --warmup 3
