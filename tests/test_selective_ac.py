"""Selective-recompute placement must reproduce the reference patterns bit-for-bit
(reference tests/test_selective_ac.py:12-64, rule at fms_fsdp/policies/ac_handler.py:43-58)."""
import pytest

from fms_fsdp_b200.models.llama import LLaMABlock
from fms_fsdp_b200.policies import apply_fsdp_checkpointing, is_checkpointed, selection_mask
from fms_fsdp_b200.policies.ac_handler import parse_fraction

T, F = True, False
CASES = [
    (0, [F] * 15), (1 / 100, [F] * 15), (-1, [F] * 15),
    (1 / 5, [F, F, T, F, F] * 3), (1 / 3, [F, T, F] * 5), (1 / 2, [T, F] * 7 + [T]),
    (3 / 5, [T, F, T, F, T] * 3), (2 / 3, [T, F, T] * 5), (1, [T] * 15), (5 / 3, [T] * 15),
    ("1/3", [F, T, F] * 5), ("1/2", [T, F] * 7 + [T]),
]


@pytest.mark.parametrize("p,expected", CASES)
def test_patterns(narrow_model_factory, p, expected):
    model = narrow_model_factory(15)
    apply_fsdp_checkpointing(model, LLaMABlock, p)
    assert [is_checkpointed(b) for b in model.layers] == expected
    assert selection_mask(15, p) == expected


def test_fraction_parser_is_not_eval():
    assert parse_fraction("1/4") == 0.25
    assert parse_fraction(0.5) == 0.5
    with pytest.raises(Exception):
        parse_fraction("__import__('os').system('true')")


def test_non_reentrant_wrapper_marks_a_single_block():
    import torch.nn as nn
    from fms_fsdp_b200.policies.ac_handler import is_checkpointed, non_reentrant_wrapper
    blk = nn.Linear(2, 2)
    assert not is_checkpointed(blk)
    assert non_reentrant_wrapper(blk) is blk and is_checkpointed(blk)


def test_selection_equals_what_the_reference_install_wraps():
    """Run the unmodified reference ``apply_fsdp_checkpointing`` (torch ``checkpoint_wrapper``) on stacks of plain blocks and
    compare WHICH blocks it wrapped with ``selection_mask``, for the depths of the model zoo and a sweep of fractions."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "fms_fsdp", "policies", "ac_handler.py")):
        pytest.skip("reference install (baseline/_ref) not present")
    fracs = ["1/7", "1/4", "1/3", "2/5", "1/2", "3/5", "2/3", "3/4", "9/10", "1"]
    depths = [10, 24, 32, 40, 48, 80]
    code = (
        "import sys, json\n"
        f"sys.path = [{ref!r}] + [p for p in sys.path if p not in ('', {root!r})]\n"
        "import importlib.util, torch.nn as nn\n"
        f"spec = importlib.util.spec_from_file_location('ref_ac', {os.path.join(ref, 'fms_fsdp', 'policies', 'ac_handler.py')!r})\n"
        "ac = importlib.util.module_from_spec(spec); spec.loader.exec_module(ac)\n"
        "from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import CheckpointWrapper\n"
        "class Blk(nn.Linear):\n    pass\n"
        "out = {}\n"
        f"for n in {depths!r}:\n"
        f"    for p in {fracs!r}:\n"
        "        m = nn.Sequential(*[Blk(2, 2) for _ in range(n)])\n"
        "        ac.apply_fsdp_checkpointing(m, Blk, p)\n"
        "        out[f'{n}:{p}'] = [isinstance(c, CheckpointWrapper) for c in m]\n"
        "print('OUT' + json.dumps(out))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/", timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    theirs = json.loads(r.stdout.split("OUT", 1)[1])
    for n in depths:
        for p in fracs:
            assert selection_mask(n, p) == theirs[f"{n}:{p}"], (n, p)
