# round 2: compute-sanitizer over the kernels added this round (k_sweep, k_global_move, k_gen_recalc, k_apply_ops_sorted,
# the rate-move kernels): memcheck, then racecheck (shared-memory hazards) on the sweep kernels
mkdir -p gpurun_out
mem() {   # tag, files, -k expression
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest $2 -k "$3" -m gpu -x -q > gpurun_out/san2_$1.log 2>&1
  echo "memcheck $1 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san2_$1.log | tail -3
}
race() {
  timeout 1700 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --launch-timeout 0 python -m pytest $2 -k "$3" -m gpu -x -q > gpurun_out/race2_$1.log 2>&1
  echo "racecheck $1 rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/race2_$1.log | sort | uniq -c | tail -6
}
mem sweep tests/test_gpu_sampler.py "sweep_routes or consistent"
mem ancient tests/test_gpu_sampler_ancient.py "consistent"
mem parity tests/test_gpu_parity.py "incremental_recalc or delta_upload or edge_lengths"
race sweep tests/test_gpu_sampler.py "sweep_routes and (hap16 or mid or tiny or ancient)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
