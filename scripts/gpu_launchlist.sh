# launch list (gpu__time_duration) of a few MCMC iterations:  bash scripts/gpu_launchlist.sh TAG CONFIG LOCI [skip] [count]
mkdir -p gpurun_out
TAG=$1; CFG=$2; L=$3; SKIP=${4:-400}; CNT=${5:-400}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file gpurun_out/${TAG}_launches.csv \
  python scripts/sampler_bench.py --config $CFG --loci $L --iterations 6 > gpurun_out/${TAG}.log 2>&1; echo "ncu list rc=$?"
