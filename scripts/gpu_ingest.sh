# ingest on the GPU box: parity tests, the drop-in chain with AlignmentProcessor.o left out, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ingest.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/ing_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/ing_pytest.log
timeout 600 python - > gpurun_out/ing_bench.log 2>&1 <<'PY'
import importlib, json, sys
sys.path.insert(0, ".")
import bench
gp = importlib.import_module("g-phocs_b200"); synth = importlib.import_module("g-phocs_b200.synth")
for cfg, L in (("dip8mig", 10000), ("hap16", 10000), ("pop6mig4", 20000)):
    print(json.dumps(bench.ingest_bench(gp, synth, 0, cfg, L)))
PY
echo "bench rc=$?"; grep "^{" gpurun_out/ing_bench.log | cut -c1-900
