mkdir -p gpurun_out
python profiles/sanitizer_workload.py 2>&1 | tail -3
(echo "# compute-sanitizer on profiles/sanitizer_workload.py (every kernel family once, small sizes), B200, round 2 (final tree)"; echo "## memcheck"; timeout 600 compute-sanitizer --tool memcheck python profiles/sanitizer_workload.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|error:" | head -8; echo "## racecheck"; timeout 900 compute-sanitizer --tool racecheck python profiles/sanitizer_workload.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard" | head -8; echo "## synccheck"; timeout 600 compute-sanitizer --tool synccheck python profiles/sanitizer_workload.py 2>&1 | grep -E "ERROR SUMMARY|Barrier error" | head -8) > gpurun_out/r02_sanitizer.txt
cat gpurun_out/r02_sanitizer.txt
