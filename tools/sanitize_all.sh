#!/bin/bash
# regenerates profiles-style sanitizer summary into gpurun_out/sanitizer_summary.txt
OUT=gpurun_out/sanitizer_summary.txt
echo "compute-sanitizer over tools/sanitize_run.py (cases: pkl, striped + hand-over, pkg = the glyph kernel routed (+ paths it hands over), pks routed, general pipeline, row band, device stroker, atlas, conic, host sink + row-packed transport, per-path status);" > $OUT
echo "command per tool: compute-sanitizer --tool T --target-processes all --print-limit 20 --log-file ... python tools/sanitize_run.py   (B200, round 2, final kernels incl. csrc/glyph_kernel.cuh)" >> $OUT
for T in memcheck racecheck initcheck synccheck; do
  echo >> $OUT; echo "==== $T" >> $OUT
  timeout 300 compute-sanitizer --tool $T --target-processes all --print-limit 20 --log-file gpurun_out/san_$T.log python tools/sanitize_run.py > gpurun_out/san_$T.out 2>&1
  grep -E "COMPUTE-SANITIZER|SUMMARY|Error|error|hazard" gpurun_out/san_$T.log | head -20 >> $OUT
  echo "---- program output" >> $OUT
  cat gpurun_out/san_$T.out >> $OUT
done
tail -5 $OUT
