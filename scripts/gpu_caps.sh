# sampler iterations/s against the loci-per-batch cap (development)
mkdir -p gpurun_out
OUT=gpurun_out/caps.log
: > $OUT
for cap in 16 8 4 2; do
  echo "cap $cap" >> $OUT
  GPHOCS_EVAL_MAX_LOCI=$cap timeout 200 python scripts/sampler_bench.py --config hap16 --loci 10000 --iterations 100 2>&1 | tail -1 | cut -c1-200 >> $OUT
  GPHOCS_EVAL_MAX_LOCI=$cap timeout 200 python scripts/sampler_bench.py --config hap16 --loci 12500 --iterations 100 2>&1 | tail -1 | cut -c1-200 >> $OUT
done
for cap in 16 8; do
  echo "cap $cap" >> $OUT
  GPHOCS_EVAL_MAX_LOCI=$cap timeout 200 python scripts/sampler_bench.py --config hap16 --loci 100000 --iterations 20 2>&1 | tail -1 | cut -c1-200 >> $OUT
  GPHOCS_EVAL_MAX_LOCI=$cap timeout 200 python scripts/sampler_bench.py --config dip8mig --loci 10000 --iterations 50 2>&1 | tail -1 | cut -c1-200 >> $OUT
done
cat $OUT
