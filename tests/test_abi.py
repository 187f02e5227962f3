"""CPU: the C-ABI library loads, exports every entry point include/elmer_b200.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest


def test_library_exports_header_symbols(b200):
    declared = b200.header_symbols()
    exported = set(b200.exported_symbols())
    assert len(declared) >= 30
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    # the Fortran shim binds exactly these names
    shim = open(os.path.join(os.path.dirname(b200.HEADER_PATH), "..", "elmerfem_b200", "fortran", "B200Solve.F90")).read()
    bound = set(re.findall(r'NAME\s*=\s*"(b200_[a-z0-9_]+)"', shim, flags=re.I))
    assert bound, "no BIND(C) names found in the Fortran shim"
    assert bound <= exported, sorted(bound - exported)


def test_no_torch_or_cuda_types_in_signatures(b200):
    text = open(b200.HEADER_PATH).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for bad in ["cudaStream", "ncclComm", "torch", "at::", "std::"]:
        assert bad not in code, bad


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "elmerfem_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".F90")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("# oracle", ""), os.path.join(dp, f)


def test_fails_loudly_without_gpu(b200):
    L = b200.lib()
    n = C.c_int(0)
    has_gpu = subprocess.call(["nvidia-smi", "-L"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0 if _which("nvidia-smi") else False
    if has_gpu:
        pytest.skip("a GPU is present")
    h = C.c_void_p(None)
    assert L.b200_create(C.byref(h)) != 0
    assert not h.value
    msg = L.b200_last_error().decode()
    assert "CUDA" in msg or "no CUDA device" in msg
    with pytest.raises(b200.B200Error):
        b200.Matrix()


def _which(x):
    import shutil
    return shutil.which(x)


def test_bench_reference_arm_contract():
    """bench.py --impl reference runs on the host alone (oracle + workload generator) and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--ne", "16", "--steps", "1", "--warmup", "3",
                          "--ref-budget", "2"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "iterations/s" and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_shim_forwards_every_keyword_the_planner_reads():
    """Every SIF keyword plan_from_sif (csrc/itersolver.cu) reads must be copied into the sif text by the Fortran shim
    (fortran/B200Solve.F90 AddStr / AddInt / AddReal / AddLog); a key that is read but not forwarded silently falls back to its default
    (IterSolve.F90:245-503 reads them from the Solver section)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "elmerfem_b200", "csrc", "itersolver.cu")).read()
    shim = open(os.path.join(root, "elmerfem_b200", "fortran", "B200Solve.F90")).read()
    read = set(re.findall(r'P\.(?:logical|real|integer|str|string|has)\(\s*"([^"]+)"', src))
    read = {k for k in read if " " in k or k.startswith("IDRS") or k.startswith("BiCG")}      # keywords, not values
    sent = {k.lower() for k in re.findall(r"CALL Add(?:Str|Int|Real|Log)\(\s*'([^']+)'", shim)}
    sent |= {k.strip().lower() for k in re.findall(r"CALL Append\(\s*'([^=']+)=", shim)}
    missing = sorted(k for k in read if k.lower() not in sent)
    assert not missing, missing
