# quick check of the migration path on the GPU box: statistics / consistency tests, then iterations/s on the two migration shapes
mkdir -p gpurun_out
TAG=${1:-migcheck}
timeout 900 python -m pytest tests/test_gpu_sampler_mig.py -x -q -k "segment_statistics or stays_consistent or uninformative" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
rm -f gpurun_out/${TAG}_bench.log
timeout 600 python scripts/sampler_bench.py --config pop6mig4 --loci 100000 --iterations 6 >> gpurun_out/${TAG}_bench.log 2>&1
timeout 300 python scripts/sampler_bench.py --config dip8mig --loci 10000 --iterations 30 >> gpurun_out/${TAG}_bench.log 2>&1
cut -c1-400 gpurun_out/${TAG}_bench.log
