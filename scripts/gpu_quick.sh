# quick dev loop on the GPU box: parity tests, short bench without extras, optional ncu (set full) of one kernel
#   bash scripts/gpu_quick.sh TAG [kernel-regex] [test files...]
mkdir -p gpurun_out
TAG=${1:-q}
KREGEX=${2:-}
shift; shift
TESTS=${@:-tests/test_gpu_parity.py tests/test_gpu_sampler.py}
timeout 900 python -m pytest $TESTS -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline --no-mcmc > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench.log
if [ -n "$KREGEX" ] && [ "$KREGEX" != "-" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mcmc > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
