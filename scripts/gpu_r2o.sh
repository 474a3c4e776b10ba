mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler_mig.py tests/test_gpu_sampler.py -x -q -k "sweep_routes or consistent or segment" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2o_pytest.log | cut -c1-300
rm -f gpurun_out/r2o_bench.log
for cfg in pop6mig4 dip8mig; do
  L=100000; [ $cfg = dip8mig ] && L=10000
  timeout 300 python scripts/sampler_bench.py --config $cfg --loci $L --iterations 10 >> gpurun_out/r2o_bench.log 2>&1
  timeout 300 python scripts/sampler_bench.py --config $cfg --loci $L --iterations 10 --stepwise >> gpurun_out/r2o_bench.log 2>&1
done
cut -c1-330 gpurun_out/r2o_bench.log
