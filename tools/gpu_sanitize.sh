OUT=gpurun_out/r2_compute_sanitizer.txt
echo "compute-sanitizer on B200, tools/sanitize_case.py (one small assemble + up to 50 GMRES iterations per kernel: fused p=3, linear tets with sixteen lanes per element, order-2 tets with a column per lane, 2-D / 3-D conv-diff, order-4 tets + chunked block SpMV, WEXPLICIT, SEXPLICIT, CG)" > $OUT
echo "--- memcheck (all nine cases)" >> $OUT
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py 2>&1 | grep -E "^ok|COMPUTE-SANITIZER|ERROR SUMMARY|Invalid|Error" | head -40 >> $OUT
echo "--- racecheck (linear tets, order-2 tets, order-4 tets + chunked SpMV, CG)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py 1 2 5 8 2>&1 | grep -E "^ok|COMPUTE-SANITIZER|RACECHECK SUMMARY|hazard|Error" | head -40 >> $OUT
echo "--- synccheck (linear tets, order-2 tets, order-4 tets, explicit types, CG)" >> $OUT
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_case.py 1 2 5 6 7 8 2>&1 | grep -E "^ok|COMPUTE-SANITIZER|ERROR SUMMARY|Error" | head -40 >> $OUT
cat $OUT
