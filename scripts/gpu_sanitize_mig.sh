# compute-sanitizer over the migration kernels (sampler_mig.cuh): memcheck, then racecheck of the per-warp shared-memory staging
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_gpu_sampler_mig.py -k "segment_statistics or stays_consistent" -m gpu -x -q > gpurun_out/san_mig.log 2>&1
echo "memcheck mig rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mig.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_gpu_sampler_mig.py -k "stays_consistent" -m gpu -x -q > gpurun_out/race_mig.log 2>&1
echo "racecheck mig rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/race_mig.log | sort | uniq -c | tail -8
