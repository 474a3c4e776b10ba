# full ncu capture of one launch of a kernel inside the sampler bench:  bash scripts/gpu_ncu_kernel.sh TAG KERNEL_REGEX CONFIG LOCI [skip]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${5:-20} -c 1 -o gpurun_out/$1 \
    python scripts/sampler_bench.py --config $3 --loci $4 --iterations 2 > gpurun_out/$1.log 2>&1; echo "ncu full $1 rc=$?"
