"""Repository contracts checkable without a GPU: where oracle/ may be imported, and the JSON line of the CPU reference
arm of bench.py."""
import ast
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_imports(path):
    """[(enclosing function or None, lineno)] of every import of the oracle package in a source file."""
    tree = ast.parse(open(path).read())
    hits = []

    def visit(node, fn):
        for child in ast.iter_child_nodes(node):
            name = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(child, ast.ImportFrom) and (child.module or "").split(".")[0] == "oracle":
                hits.append((fn, child.lineno))
            if isinstance(child, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in child.names):
                hits.append((fn, child.lineno))
            visit(child, name)

    visit(tree, None)
    return hits


def test_product_package_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "text2nerf_b200", "**", "*.py"), recursive=True):
        assert _oracle_imports(path) == [], path


def test_bench_touches_the_oracle_only_in_its_cpu_legs():
    hits = _oracle_imports(os.path.join(ROOT, "bench.py"))
    assert hits, "the cpu_baseline / --impl reference legs run the oracle"
    assert {fn for fn, _ in hits} <= {"oracle_spec", "cpu_reference_leg", "cpu_reference_regrad", "cpu_reference_kinks"}, hits


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    # "reference" when oracle/_ref (the unmodified reference modules, oracle/make_ref.py) is present, else the port
    want = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")) else "port"
    assert line["cpu_baseline"]["kind"] == want and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: bench.py's own arm stops with a clear message when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr and out.stdout.strip() == ""


def test_oracle_ref_is_the_unmodified_reference():
    """oracle/_ref (when built) holds byte-identical copies of the three reference modules of the path, and stays out of
    the git history."""
    import hashlib
    man = os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")
    if not os.path.exists(man):
        import pytest
        pytest.skip("oracle/_ref not built (no reference checkout at build time)")
    m = json.load(open(man))
    assert set(m["files"]) == {"models/tensorBase.py", "models/tensoRF.py", "models/sh.py"}
    for rel, sha in m["files"].items():
        assert hashlib.sha256(open(os.path.join(ROOT, "oracle", "_ref", rel), "rb").read()).hexdigest() == sha
        src = os.path.join(m["source"], rel)
        if os.path.exists(src):
            assert hashlib.sha256(open(src, "rb").read()).hexdigest() == sha
    ign = subprocess.run(["git", "check-ignore", "oracle/_ref/models/sh.py"], capture_output=True, text=True, cwd=ROOT)
    assert ign.returncode == 0
