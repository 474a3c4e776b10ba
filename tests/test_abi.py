"""CPU: the C-ABI library loads and exports every symbol include/gphocs_b200.h declares
(no compute calls — there is no GPU here), and refuses to work without a device."""
import ctypes
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gp = importlib.import_module("g-phocs_b200")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "gphocs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    # type names in front of a function-pointer declarator "( *fn)(" are not symbols
    return sorted({n for n in names if n not in ("defined", "int", "void", "double", "long")})


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(gp.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return gp.lib()


def test_exports_every_declared_symbol(built):
    out = subprocess.run(["nm", "-D", "--defined-only", gp.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = _header_symbols()
    assert len(declared) >= 50
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    # and the Python binding declares a signature for each of them
    assert sorted(built._declared) == declared


def test_reference_call_surface_is_complete(built):
    ref_api = ["createLocusData", "initializeLocusData", "freeLocusData", "attachLeaf_UNUSED", "setLocusMutationRate",
               "getLocusMutationRate", "computeAllConditionals", "computeLocusDataLikelihood", "computePatternLogLikelihood",
               "computeLocusDataLikelihood_deb", "addSitePatterns", "reduceSitePatterns", "checkLocusDataLikelihood",
               "revertToSaved", "resetSaved", "adjustGenNodeAge", "scaleAllNodeAges", "executeGenSPR",
               "copyGenericTreeToLocus", "printLocusGenTree", "printLocusDataStats", "printLocusDataPatterns",
               "computePairwiseLCAs", "getSortedAges", "getLocusDataLikelihood", "getLocusRoot", "getNodeAge",
               "getNodeFather", "getNodeSon"]
    for name in ref_api:   # LocusDataLikelihood.h:54-341
        assert hasattr(built, name), name


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    with pytest.raises(RuntimeError):
        gp.LociStore(3, [0, 1], [0, 1], np.frombuffer(b"TCT", np.uint8).reshape(1, 3), [1], [5])


def test_product_does_not_link_the_oracle(built):
    out = subprocess.run(["ldd", gp.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "gphocs_ref" not in out
    src_dir = os.path.join(ROOT, "g-phocs_b200")
    for dirpath, _, files in os.walk(src_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f
