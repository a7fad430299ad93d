#!/bin/bash
# run 2: exactness after canonical-order rework, bench with CPU thread probing, ncu launch list + full capture (round-1 v1 kernels)
mkdir -p gpurun_out
(nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os; print(len(os.sched_getaffinity(0)))") > gpurun_out/host2.txt 2>&1
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run build 600 python __graft_entry__.py
run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider
TAILN=30 run bench 1500 python bench.py --steps 2 --warmup 3
K='regex:matvec_kernel|attn_kernel|gemm_f16|attention_f32|layernorm|embed_kernel|finalize|im2col|cls_row'
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 2000 -c 700 --csv --log-file gpurun_out/launches_r1v1.csv python bench.py --steps 1 --warmup 1 --no-cpu
run ncu_full 600 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 400 -c 5 -o gpurun_out/prof_matvec_r1v1 python bench.py --steps 1 --warmup 1 --no-cpu
ls -la gpurun_out
