#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run build 600 python __graft_entry__.py
run e2e 600 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "not gemm and not attention_matches and not layernorm"
TAILN=30 run trace 900 python tools/mega_trace.py
