#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for TOOL in memcheck; do
  ( timeout 900 compute-sanitizer --tool $TOOL --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -k "large_windows and (1-pd0 or 1-pd4 or 2-pd5 or 2-pd2)" ) > gpurun_out/r2bh_san_${TOOL}_attn_$i.log 2>&1
  echo "$TOOL attn run $i: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2bh_san_${TOOL}_attn_$i.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2bh_san_${TOOL}_attn_$i.log | tail -1)"
done
done
( timeout 600 compute-sanitizer --tool synccheck --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -k "large_windows and (1-pd0)" ) > gpurun_out/r2bh_san_synccheck.log 2>&1; grep -E "ERROR SUMMARY| passed| failed" gpurun_out/r2bh_san_synccheck.log | tail -2
( timeout 600 compute-sanitizer --tool initcheck --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -k "large_windows and (1-pd0)" ) > gpurun_out/r2bh_san_initcheck.log 2>&1; grep -E "ERROR SUMMARY| passed| failed" gpurun_out/r2bh_san_initcheck.log | tail -2
MICFORMER_ATTN_BWD_SIMT=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -k "large_windows and (1-pd0)" > gpurun_out/r2bh_san_memcheck_simt.log 2>&1; echo "memcheck with SIMT backward: $(grep -E 'ERROR SUMMARY| passed| failed' gpurun_out/r2bh_san_memcheck_simt.log | tail -2 | tr '\n' ' ')"
