# k_eval timing of the library variants under gpurun_variants/ (development)
mkdir -p gpurun_out
OUT=gpurun_out/variants.log
: > $OUT
run() { timeout 120 python scripts/keval_variants.py $1 ${2:-pop6mig4} ${3:-100000} 2>&1 | tail -1 >> $OUT; }
run g-phocs_b200/csrc/libgphocs_b200.so
GPHOCS_EVAL_PREFETCH=0 run g-phocs_b200/csrc/libgphocs_b200.so
for v in sel lf sellf; do run gpurun_variants/lib_$v.so; done
run g-phocs_b200/csrc/libgphocs_b200.so hap16
run gpurun_variants/lib_sellf.so hap16
cat $OUT
