mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log; tail -5 gpurun_out/r1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1_smoke.log; tail -3 gpurun_out/r1_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1_bench.log; tail -3 gpurun_out/r1_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3 -c 2 -o gpurun_out/r1_prof_keval python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full rc=$?"
