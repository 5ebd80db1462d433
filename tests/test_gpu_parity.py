"""pytest -m gpu: parity of the CUDA path (through the C ABI) against the CPU oracle and the
reference-generated golden fixtures.  The checks live in tests/gpu_checks.py."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _names():
    import ast
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'gpu_checks.py')).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', None) == 'CHECKS':
            return [k.value for k in node.value.keys]
    return []


@pytest.mark.parametrize('name', _names())
def test_gpu_check(name):
    import torch
    assert torch.cuda.is_available(), 'pytest -m gpu needs a GPU'
    import gpu_checks
    res = gpu_checks.CHECKS[name]()
    torch.cuda.synchronize()
    print(name, res)
