"""CPU: the host runtime inside the product library — thread teams and the fiber scheduler that stands in for
libgomp when the reference host is linked against libgphocs_b200.so (g-phocs_b200/csrc/host_runtime.cpp)."""
import importlib
import subprocess

import pytest

gp = importlib.import_module("g-phocs_b200")


@pytest.mark.parametrize("fibers,parks,threads", [(1, 1, 1), (7, 3, 2), (1000, 5, 4), (40000, 2, 8), (333, 0, 3)])
def test_fiber_regions_run_every_iteration_once(fibers, parks, threads):
    lib = gp.lib()
    want = parks * fibers * (fibers + 1) // 2
    assert lib.gphocsFiberSelfTest(fibers, parks, threads, 1) == want
    # the same region on plain OS threads (direct mode: parks are no-ops)
    assert lib.gphocsFiberSelfTest(fibers, parks, threads, 0) == want


def test_library_exports_the_openmp_entry_points_and_needs_no_libgomp():
    out = subprocess.run(["nm", "-D", "--defined-only", gp.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for sym in ("GOMP_parallel", "omp_get_thread_num", "omp_get_num_threads", "omp_get_max_threads", "omp_set_num_threads"):
        assert f" T {sym}" in out, sym
    assert "libgomp" not in subprocess.run(["ldd", gp.LIB_PATH], capture_output=True, text=True).stdout
