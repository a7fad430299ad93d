#!/bin/bash
# Full 1-GPU check-out.   gpurun --timeout 2400 -- 'bash tools/gpu_full.sh'
mkdir -p gpurun_out
TAG=${TAG:-r2_v6}
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=6 run canary6 300 python tools/canary.py || exit 1
TAILN=8 run pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider
TAILN=3 run bench_q5k 700 python bench.py --wtype q5_k --tokens 256 --steps 2 --warmup 1
TAILN=3 run bench 600 python bench.py
echo done
