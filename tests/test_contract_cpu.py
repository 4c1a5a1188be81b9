"""CPU checks of the repository contract: the reference arm of bench.py prints a well-formed JSON line without a GPU, and
the product (library sources + host package) never touches the oracle (the oracle is test infrastructure only)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "clima-oceananigans.jl_b200")


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("grid-point updates/sec") and d["unit"] == "grid-point updates/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_product_never_uses_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle_cpu|liboracle", re.M)
    offenders = []
    for base, _, files in os.walk(PKG):
        if os.path.basename(base) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if pat.search(text):
                    offenders.append(os.path.join(base, f))
    assert not offenders, offenders
    assert not pat.search(open(os.path.join(ROOT, "include", "ocean_b200.h")).read())
