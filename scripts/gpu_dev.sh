# dev loop on the GPU box: parity tests, a short bench, ncu of k_eval (full + incremental launches)
mkdir -p gpurun_out
TAG=${1:-dev}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.log
if [ "${2:-}" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3 -c 1 -o gpurun_out/${TAG}_prof_keval python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
