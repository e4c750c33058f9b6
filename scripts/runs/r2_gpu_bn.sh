#!/bin/bash
# compute-sanitizer on the kernels added last: row-reuse conv weight gradient (cp.async zero fill), unpatch-view GEMMs (5-D TMA)
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  ( timeout 900 compute-sanitizer --tool $TOOL --print-limit 3 python -m pytest tests/test_gpu_tc.py -q -k "(conv3 and (24-0-8 or 40-24-8 or 16-0-8)) or unpatch_view" ) > gpurun_out/r2bn_san_${TOOL}.log 2>&1
  echo "$TOOL: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2bn_san_${TOOL}.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2bn_san_${TOOL}.log | tail -1)"
done
