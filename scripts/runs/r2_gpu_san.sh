#!/bin/bash
# compute-sanitizer on the tensor-core kernels at small shapes; logs -> profiles/r02_sanitizer_*.txt
mkdir -p gpurun_out
run() { name=$1; shift; ( timeout 900 compute-sanitizer --tool $TOOL --print-limit 3 "$@" ) > gpurun_out/r2san_${TOOL}_$name.log 2>&1; echo "$TOOL $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2san_${TOOL}_$name.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2san_${TOOL}_$name.log | tail -1)"; }
TOOL=memcheck
run fused python -m pytest tests/test_gpu_fused.py -q -x -k "(48-128 or 24-4096 or 64-130 or 192-200 or heads0 or dims0 or 2-dims2 or 2-dims3) and not blocks"
run gemm python -m pytest tests/test_gpu_tc.py -q -x -k "epilogues or linear"
run conv python -m pytest tests/test_gpu_tc.py -q -x -k "conv3 and (32-0-16 or 24-24-16 or 16-0-8)"
run attn python -m pytest tests/test_gpu_tc.py -q -x -k "large_windows and (B0 or B4 or B5)"
TOOL=racecheck
run fused python -m pytest tests/test_gpu_fused.py -q -x -k "(48-128 or 64-130 or heads0 or 2-dims3) and not blocks and not 65536"
run gemm python -m pytest tests/test_gpu_tc.py -q -x -k "epilogues"
run conv python -m pytest tests/test_gpu_tc.py -q -x -k "conv3 and 16-0-8"
