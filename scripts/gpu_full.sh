# full single-GPU check: all GPU tests, smoke, both bench arms with default flags, ncu launch list + full profile
mkdir -p gpurun_out
TAG=${1:-full}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/${TAG}_bench_ref.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eval|k_gen_eval' -s 6 -c 2 -o gpurun_out/${TAG}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
# sampler and ingest kernels: launch list of three MCMC iterations, full profile of one launch each
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 360 --csv --log-file gpurun_out/${TAG}_sampler_launches.csv python scripts/sampler_bench.py --config hap16 --loci 100000 --iterations 3 > gpurun_out/${TAG}_ncu_sampler_launches.log 2>&1; echo "ncu sampler launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_smp_spr_propose|k_smp_age_propose|k_smp_tau_propose' -s 30 -c 3 -o gpurun_out/${TAG}_prof_sampler python scripts/sampler_bench.py --config hap16 --loci 100000 --iterations 2 > gpurun_out/${TAG}_ncu_sampler.log 2>&1; echo "ncu sampler rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ingest -c 4 -o gpurun_out/${TAG}_prof_ingest python -c "
import importlib, sys, tempfile, os
sys.path.insert(0, '.')
gp = importlib.import_module('g-phocs_b200'); synth = importlib.import_module('g-phocs_b200.synth')
m = synth.config('dip8mig'); tmp = tempfile.mkdtemp(); p = os.path.join(tmp, 's.txt')
synth.generate(m, 10000, seed=4242, seqfile=p)
a = gp.Alignment.read(p, synth.sample_slots(m)); print(a.P, a.U); a.close()
" > gpurun_out/${TAG}_ncu_ingest.log 2>&1; echo "ncu ingest rc=$?"
