# compute-sanitizer racecheck (shared-memory hazards) over the kernels that coordinate through shared memory
mkdir -p gpurun_out
run() {   # tag, file, -k expression
  timeout 1700 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --launch-timeout 0 python -m pytest "$2" -k "$3" -m gpu -x -q > gpurun_out/race_$1.log 2>&1
  echo "$1 rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/race_$1.log | sort | uniq -c | tail -6
}
run ingest tests/test_gpu_ingest.py "golden"
run parity tests/test_gpu_parity.py "golden or rounds"
run sampler tests/test_gpu_sampler.py "state_stays or routes"
run sampler_mig tests/test_gpu_sampler_mig.py "consistent"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
