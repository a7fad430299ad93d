#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run build 600 python __graft_entry__.py
run e2e 600 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider -x
TAILN=30 run trace 900 python tools/mega_trace.py
TAILN=30 run bench 1500 python bench.py --steps 2 --warmup 3
run ncu_vision 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_f16|attention_f32|layernorm|im2col|cls_row' -c 460 --csv --log-file gpurun_out/launches_vision_r1v1.csv python bench.py --steps 1 --warmup 1 --no-cpu
