#!/bin/bash
# first GPU validation pass: every stage in its own process with its own timeout (a hung kernel must not eat the box)
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
(nproc; lscpu | head -25; free -g) > gpurun_out/host.txt 2>&1
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 10 $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run build 600 python __graft_entry__.py
run k_matvec 300 python -m pytest tests/test_kernels_gpu.py -q -k "matvec" -p no:cacheprovider
run k_ln_attn 300 python -m pytest tests/test_kernels_gpu.py -q -k "layernorm or attention" -p no:cacheprovider
run k_gemm 300 python -m pytest tests/test_kernels_gpu.py -q -k "gemm" -p no:cacheprovider
run e2e 900 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider
run smoke 600 python __graft_entry__.py --smoke
TAILN=40 run bench 1500 python bench.py --steps 2 --warmup 3
