# round 2, first GPU visit: the one-launch sweep against the stepwise route (bit-identical chains), then iterations/s
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py -x -q -k "sweep_routes or stays_consistent" > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest.log
for L in 100000 12500 10000; do
  timeout 300 python scripts/sampler_bench.py --config hap16 --loci $L --iterations 30 >> gpurun_out/r2a_bench.log 2>&1
  timeout 300 python scripts/sampler_bench.py --config hap16 --loci $L --iterations 30 --stepwise >> gpurun_out/r2a_bench.log 2>&1
done
cat gpurun_out/r2a_bench.log
