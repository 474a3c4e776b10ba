# full single-GPU check: all GPU tests, smoke, both bench arms with default flags, ncu launch list + full profile
mkdir -p gpurun_out
TAG=${1:-full}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/${TAG}_bench_ref.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eval|k_gen_eval' -s 6 -c 2 -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
