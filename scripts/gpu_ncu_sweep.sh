# full ncu capture of one k_sweep launch (hap16, 100k loci); TAG = output name under gpurun_out/
mkdir -p gpurun_out
TAG=${1:-sweep}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -o gpurun_out/$TAG \
    python scripts/sampler_bench.py --config ${2:-hap16} --loci ${3:-100000} --iterations 2 > gpurun_out/${TAG}.log 2>&1; echo "ncu full rc=$?"
