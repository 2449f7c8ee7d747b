# round 1, last session: new pieces (aggregation "min", N-ImageNet upstream wrappers, background-activity filter, filter objects)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_filters.py tests/test_gpu_nimagenet.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_new.log 2>&1; echo pytest rc=$?
tail -25 gpurun_out/pytest_new.log
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_filters.py -m gpu -q -k "background_activity_matches and tiny or argument_checks" > gpurun_out/memcheck_ba.log 2>&1; echo memcheck rc=$?
tail -6 gpurun_out/memcheck_ba.log
timeout 200 python bench_extra.py --only filters > gpurun_out/bench_filters.log 2>&1; echo bench rc=$?
cat gpurun_out/bench_filters.log
