"""The product (scanobjectnn_b200/, bench.py's own arm) must never route through the checker: oracle/ may only be imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Static checks, CPU."""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, node.lineno
        elif isinstance(node, ast.ImportFrom):
            yield (node.module or ""), node.lineno


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scanobjectnn_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(dirpath, f)
                bad += [(p, ln) for mod, ln in _imports(p) if mod == "oracle" or mod.startswith("oracle.")]
    assert not bad, bad
    # and no native source of the product includes anything from oracle/
    for f in os.listdir(os.path.join(pkg, "csrc")):
        if f.endswith((".cu", ".cuh")):
            assert not re.search(r'#include\s+"[^"]*oracle', open(os.path.join(pkg, "csrc", f)).read()), f


def test_bench_uses_the_oracle_only_in_the_baseline_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    top_level = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    for n in top_level:
        mods = [a.name for a in n.names] if isinstance(n, ast.Import) else [n.module or ""]
        assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), "bench.py imports oracle at module level"
    # every function that imports the oracle is one of the baseline legs (methods are named by their class: CpuForward.__init__)
    allowed = {"cpu_forward_rate", "run_reference_arm", "reference_arm", "cpu_baseline"}
    owners = {}
    for cls in [n for n in ast.walk(tree) if isinstance(n, ast.ClassDef)]:
        for fn in [n for n in cls.body if isinstance(n, ast.FunctionDef)]:
            owners[id(fn)] = cls.name
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = any((isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")) or
                   (isinstance(n, ast.Import) and any(a.name.startswith("oracle") for a in n.names)) for n in ast.walk(fn))
        if uses:
            name = (owners.get(id(fn), "") + "." + fn.name).lower()
            assert fn.name in allowed or "cpu" in name or "reference" in name, name


def test_library_loader_has_no_fallback():
    src = open(os.path.join(ROOT, "scanobjectnn_b200", "_lib.py")).read()
    assert "libpsa.so" in src and "raise" in src          # missing library -> loud failure (exercised in tests/test_abi.py)
