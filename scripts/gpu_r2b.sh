# launch lists of an MCMC iteration (one-launch sweep route) at 100k and 12.5k loci + a full ncu capture of k_sweep
mkdir -p gpurun_out
for L in 100000 12500; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2b_launches_$L.csv \
    python scripts/sampler_bench.py --config hap16 --loci $L --iterations 4 > gpurun_out/r2b_l$L.log 2>&1; echo "ncu list $L rc=$?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -o gpurun_out/r2b_sweep \
    python scripts/sampler_bench.py --config hap16 --loci 100000 --iterations 2 > gpurun_out/r2b_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -5
