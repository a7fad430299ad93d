#!/bin/bash
# First GPU-box call of round 2 (≈ 6 min): re-validate HEAD, try the prepared-but-never-run pieces, re-measure.
#   gpurun --timeout 700 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary 240 python tools/canary.py || { echo "CANARY FAILED - aborting"; exit 1; }
TAILN=4 run pytest_gpu 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider
# experimental flag-in-data megakernel: a protocol bug is a hang, so everything runs under a short timeout
MINIGPT4_B200_MEGA_LL=1 TAILN=6 run canary_ll 150 python tools/canary.py
MG4_EXPERIMENTAL=1 TAILN=6 run pytest_experimental 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q -p no:cacheprovider
NOTRACE=1 TAILN=1 run ab_default 120 python tools/mega_trace.py
[ -s gpurun_out/canary_ll.log ] && grep -q MISMATCH gpurun_out/canary_ll.log || NOTRACE=1 MINIGPT4_B200_MEGA_LL=1 TAILN=1 run ab_ll 120 python tools/mega_trace.py
TAILN=22 run trace 200 python tools/mega_trace.py
MINIGPT4_B200_MEGA_LL=1 TAILN=22 run trace_ll 200 python tools/mega_trace.py   # "barrier" column = waiting for the input data
# experimental token-split vision GEMMs: correctness first (encode + kernel seams), then encode time
MINIGPT4_B200_VISION_TSPLIT=1 TAILN=4 run pytest_tsplit 300 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "encode_image or chat_flow or gemm"
TAILN=1 run encode_default 200 python tools/prof_vision.py
MINIGPT4_B200_VISION_TSPLIT=1 TAILN=1 run encode_tsplit 200 python tools/prof_vision.py
MINIGPT4_B200_VISION_TSPLIT=1 MINIGPT4_B200_VISION_SPLITK=3 TAILN=1 run encode_splitk 200 python tools/prof_vision.py
TAILN=3 run bench 600 python bench.py --steps 3 --warmup 3 --no-cpu
TAILN=8 run crosscheck_vllm 400 python tools/crosscheck_vllm_gguf.py
echo done
