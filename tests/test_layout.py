"""CPU: repository rules -- the product never touches the oracle or a CPU fallback."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".js")):
                yield os.path.join(base, f)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(import|from)\s+oracle\b|libmtz_oracle|#\s*include\s*[\"<].*oracle|dlopen.*oracle", re.M)
    bad = [p for p in _py_files(os.path.join(ROOT, "manatee_b200")) if pat.search(open(p).read())]
    assert not bad, bad


def test_oracle_files_say_they_are_test_infrastructure():
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith((".c", ".h", ".py")):
            head = open(os.path.join(ROOT, "oracle", f)).read(1500)
            assert "TEST INFRASTRUCTURE" in head.upper() or "test infrastructure" in head, f


def test_no_reference_sources_copied_and_layout_present():
    for d in ("tests", "tests/golden", "oracle", "include", "profiles", "manatee_b200/csrc"):
        assert os.path.isdir(os.path.join(ROOT, d)), d
    for f in ("bench.py", "__graft_entry__.py", "include/manatee_gpu.h"):
        assert os.path.isfile(os.path.join(ROOT, f)), f
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in gi and "*.so" in gi
