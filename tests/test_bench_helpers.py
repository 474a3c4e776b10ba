"""CPU: the parts of bench.py that do not need a GPU — the byte formulas of the roofline, the parsing of
the nvidia-smi clock samples, the config object both arms print, and the refusal to run the product arm without a CUDA
device (there is no CPU path to fall back to)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_byte_formulas_of_the_roofline():
    n, P, E, Q, B = 24, np.array([10.0, 23.0]), np.array([49.0]), 11, 4
    # 32*P*(2n-1) + 24*(2n-1) + 8*P + 8 per locus
    assert np.array_equal(bench.algorithmic_bytes_data(n, P), 32 * P * 47 + 24 * 47 + 8 * P + 8)
    # this layout: internal vectors written once, leaves as 4-bit masks (one 64-bit word per 16 leaves and column)
    assert np.array_equal(bench.layout_bytes_data(n, P), 32 * P * 23 + 8 * P * 2 + 8 * P + 16 * 47 + 24)
    assert np.all(bench.layout_bytes_data(n, P) < bench.algorithmic_bytes_data(n, P))
    assert np.array_equal(bench.algorithmic_bytes_gen(E, Q, B), 32 * E + 16 * 15 + 8)


def test_clock_samples_are_parsed_and_slowdowns_are_named():
    cs = bench.ClockSampler(0)
    cs.proc = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(30)"])
    cs.rows = ["0, 1965, 1965, 612.3, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
               "0, 1950, 1965, 701.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
               "0, 1800, 1965, 690.0, 0x0000000000000040, Not Active, Active, Not Active, Not Active",
               "garbage line", "0, [N/A], 1965, 1, 0, Not Active, Not Active, Not Active, Not Active"]
    out = cs.stop()
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]


def test_both_arms_describe_the_same_workload():
    a = bench.common_config("pop6mig4", 100_000, 2)
    assert a["loci_per_gpu"] == 100_000 and a["n_gpus"] == 2 and a["host_cores"] == os.cpu_count()
    assert "24 leaves" in a["workload"] and "4 migration bands" in a["workload"] and "model" not in a
    json.dumps(a)


def test_product_arm_refuses_to_run_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-mcmc", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in (r.stdout + r.stderr)
