# 2 GPUs: the two-rank sampler tests (incl. cross-rank locus rates), then the whole GPU suite on one of them, then bench at N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sampler_multi.py -x -q > gpurun_out/r2i_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/r2i_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err; echo "bench2 rc=$?"; tail -c 400 gpurun_out/r2i_bench2.err
