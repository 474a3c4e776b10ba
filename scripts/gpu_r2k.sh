mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "delta_upload or incremental_recalc" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-mcmc --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/r2k_bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[-1])
print(json.dumps(b['e2e'])[:900]); print(b['value'], b['ms_per_step'])
PY
