mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sampler_mig.py -x -q -k "consistent or segment or prior or (posterior and dip8mig)" --durations=5 > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2w_pytest.log | cut -c1-300
timeout 300 python scripts/sampler_bench.py --config pop6mig4 --loci 100000 --iterations 10 2>&1 | cut -c1-240
timeout 300 python scripts/sampler_bench.py --config dip8mig --loci 10000 --iterations 20 2>&1 | cut -c1-240
