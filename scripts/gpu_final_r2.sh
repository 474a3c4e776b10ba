# round 2, final measurements on one GPU: the bench line as the driver runs it, the reference arm, launch list and ncu
# captures of the migration kernels after their second pass
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/h_bench_reference_arm.json 2> gpurun_out/h_ref.err; echo "ref rc=$?"
bash scripts/gpu_launchlist.sh h_pop6mig4_100k pop6mig4 100000 400 400
bash scripts/gpu_ncu_kernel.sh h_smg_spr k_smg_spr_propose pop6mig4 100000 20
bash scripts/gpu_ncu_kernel.sh h_smg_age k_smg_age_propose pop6mig4 100000 10
cut -c1-600 gpurun_out/h_bench_n1.json
