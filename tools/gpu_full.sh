#!/bin/bash
# One comprehensive GPU-box run: build check, GPU parity tests, both bench arms, ncu launch list + full captures, megakernel trace.
# Every leg is bounded by its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-r1}
run() { name=$1; shift; t=$1; shift; echo "=== $name"; timeout -k 5 $t "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
nproc > gpurun_out/nproc.log; cat /sys/fs/cgroup/cpu.max >> gpurun_out/nproc.log 2>/dev/null
run build 300 python __graft_entry__.py || exit 1
run canary 280 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider -x -k "megakernel" || { echo "CANARY FAILED - aborting"; exit 1; }
[ -n "$SKIP_TESTS" ] || TAILN=6 run pytest_gpu 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider
TAILN=3 run bench 900 python bench.py
[ -n "$SKIP_REF" ] || TAILN=3 run bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
TAILN=22 run trace 300 python tools/mega_trace.py
NOTRACE=1 TAILN=2 run notrace 300 python tools/mega_trace.py
if [ -z "$SKIP_NCU" ]; then
KREG='regex:decode_megakernel|gemm_f16|attention_f32|layernorm|matvec_kernel|attn_kernel|stage_kernel|im2col|cls_row|embed_kernel|finalize_kernel|add_kernel'
TAILN=3 run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 14000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu
TAILN=3 run ncu_mega 600 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 20 -c 2 -o gpurun_out/${TAG}_mega -f python bench.py --steps 1 --warmup 1 --no-cpu
TAILN=3 run ncu_vision 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_f16|attention_f32|layernorm' -s 20 -c 12 -o gpurun_out/${TAG}_vision -f python bench.py --steps 1 --warmup 1 --no-cpu
fi
echo done
