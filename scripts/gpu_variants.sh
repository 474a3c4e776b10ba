# k_eval timing of the library variants under gpurun_variants/ (development)
mkdir -p gpurun_out
OUT=gpurun_out/variants.log
: > $OUT
run() { GPHOCS_EVAL_SMEM_BUDGET=$2 timeout 120 python scripts/keval_variants.py $1 ${3:-pop6mig4} ${4:-100000} 2>&1 | tail -1 >> $OUT; }
run g-phocs_b200/csrc/libgphocs_b200.so 65536
run gpurun_variants/lib_k3b0.so 65536
run gpurun_variants/lib_k3b9.so 65536
run gpurun_variants/lib_k2b0.so 18400
run gpurun_variants/lib_k2b10.so 22300
run gpurun_variants/lib_k2b10.so 20200
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv >> $OUT
cat $OUT
