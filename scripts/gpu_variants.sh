# k_eval timing of the library variants under gpurun_variants/ (development)
mkdir -p gpurun_out
OUT=gpurun_out/variants.log
: > $OUT
run() { timeout 120 python scripts/keval_variants.py $1 ${2:-pop6mig4} ${3:-100000} 2>&1 | tail -1 >> $OUT; }
for lib in g-phocs_b200/csrc/libgphocs_b200.so gpurun_variants/lib_*.so; do for cfg in pop6mig4 hap16; do run $lib $cfg; done; done
cat $OUT
