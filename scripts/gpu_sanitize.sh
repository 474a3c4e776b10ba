# compute-sanitizer passes over the newer kernels (ingest, sampler incl. migration / ancient / whole-sweep paths)
mkdir -p gpurun_out
run() {   # tag, file, -k expression
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest "$2" -k "$3" -m gpu -x -q > gpurun_out/san_$1.log 2>&1
  echo "$1 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_$1.log | tail -3
}
run ingest tests/test_gpu_ingest.py "golden or random or phasing or malformed"
run sampler tests/test_gpu_sampler.py "consistent or fused"
run sampler_mig tests/test_gpu_sampler_mig.py "consistent or segment"
run sampler_ancient tests/test_gpu_sampler_ancient.py "consistent or trace"
