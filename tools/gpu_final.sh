#!/bin/bash
# Round-end GPU-box run: canary, the GPU tests touched this session, full bench (both arms), ncu captures of the final kernels.
mkdir -p gpurun_out
TAG=${TAG:-r1_v4}
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=4 run canary 240 python tools/canary.py || { echo "CANARY FAILED - aborting"; exit 1; }
TAILN=6 run pytest_vision 500 python -m pytest tests/test_kernels_gpu.py tests/test_quantized_vision.py tests/test_e2e_gpu.py -m gpu -x -q -p no:cacheprovider -k "layernorm or attention or quantised or expansion or encode_image or chat_flow or gemm"
TAILN=3 run bench 700 python bench.py
TAILN=2 run ncu_mega 420 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 10 -c 2 -o gpurun_out/${TAG}_mega -f python tools/prof_decode.py
TAILN=2 run ncu_vision 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_f16|attention_f32|layernorm|im2col|cls_row' -s 470 -c 470 --csv --log-file gpurun_out/${TAG}_vision_launches.csv python tools/prof_vision.py
TAILN=3 run bench_ref 400 python bench.py --impl reference --steps 2 --warmup 1
echo done
