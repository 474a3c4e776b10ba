# quick check of the one-launch sweep on the GPU box: route equivalence tests, iterations/s at 100k and 12.5k loci, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py -x -q -k "sweep_routes or stays_consistent" > gpurun_out/sweepcheck_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/sweepcheck_pytest.log
rm -f gpurun_out/sweepcheck_bench.log
for L in 100000 12500; do
  timeout 300 python scripts/sampler_bench.py --config hap16 --loci $L --iterations 30 >> gpurun_out/sweepcheck_bench.log 2>&1
done
cut -c1-330 gpurun_out/sweepcheck_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 20 --csv --log-file gpurun_out/sweepcheck_launches.csv \
    python scripts/sampler_bench.py --config hap16 --loci 100000 --iterations 4 > /dev/null 2>&1; grep k_sweep gpurun_out/sweepcheck_launches.csv | head -2
