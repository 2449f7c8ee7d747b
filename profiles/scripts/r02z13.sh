mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool initcheck python profiles/sanitizer_workload.py > gpurun_out/initcheck.log 2>&1; echo rc=$?
grep -E "ERROR SUMMARY|Uninitialized" gpurun_out/initcheck.log | sort | uniq -c | head -10
grep -A12 "Uninitialized" gpurun_out/initcheck.log | grep -E "at .*evrep|at .*k_" | sort | uniq -c | sort -rn | head -20
