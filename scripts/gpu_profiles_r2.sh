# round 2 profile batch (one GPU): launch lists with gpu__time_duration, full ncu captures of the dominant kernels
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc
# the bench step (k_eval + k_gen_eval + k_gen_reduce + k_reduce_sum) and the e2e routes: launch list of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/p2_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-mcmc > gpurun_out/p2_ncu_bench.log 2>&1; echo "ncu bench list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^k_eval|k_gen_eval' -s 6 -c 2 -o gpurun_out/p2_keval \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mcmc > gpurun_out/p2_ncu_keval.log 2>&1; echo "ncu k_eval full rc=$?"
# MCMC iterations: launch lists
bash scripts/gpu_launchlist.sh p2_hap16_100k hap16 100000 60 60
bash scripts/gpu_launchlist.sh p2_hap16_12k hap16 12500 60 60
bash scripts/gpu_launchlist.sh p2_pop6mig4_100k pop6mig4 100000 700 400
bash scripts/gpu_launchlist.sh p2_ancient_50k ancient 50000 100 80
# full captures: the sweep kernel and the global-move kernel
bash scripts/gpu_ncu_kernel.sh p2_k_sweep k_sweep hap16 100000 3
bash scripts/gpu_ncu_kernel.sh p2_k_global_move k_global_move hap16 100000 9
ls -la gpurun_out | grep p2_
