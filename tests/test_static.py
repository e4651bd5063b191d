"""Static checks of the modules that only run on a GPU box: a typo there would otherwise surface after the round."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("path", ["tests/test_gpu_parity.py", "tests/test_gpu_distributed.py", "bench.py", "scripts/dist_check.py",
                                  "scripts/readback_bench.py", "__graft_entry__.py", "damavand_b200/circuit.py"])
def test_every_name_resolves(path):
    tree = ast.parse(open(os.path.join(ROOT, path)).read())
    defined = set(dir(builtins)) | {"__file__", "__name__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                defined.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            defined.add(node.name)
        elif isinstance(node, ast.arg):
            defined.add(node.arg)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, ast.Store):
            defined.add(node.id)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            defined.add(node.name)
    missing = sorted({n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined})
    assert missing == [], f"{path}: undefined names {missing}"
