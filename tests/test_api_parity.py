"""Every public top-level name (and public method) of the reference's modules must exist under the same import path
(`fms_fsdp.*` aliases to `fms_fsdp_b200.*`).  The reference tree is read-only input at /root/reference; the test is
skipped when it is not mounted (e.g. in CI)."""
import ast
import importlib
import os

import pytest

REF = "/root/reference"
MODULES = {
    "fms_fsdp/config/training.py": "fms_fsdp.config.training",
    "fms_fsdp/utils/config_utils.py": "fms_fsdp.utils.config_utils",
    "fms_fsdp/utils/train_utils.py": "fms_fsdp.utils.train_utils",
    "fms_fsdp/utils/checkpointing_utils.py": "fms_fsdp.utils.checkpointing_utils",
    "fms_fsdp/utils/dataloader_utils.py": "fms_fsdp.utils.dataloader_utils",
    "fms_fsdp/utils/dataset_utils.py": "fms_fsdp.utils.dataset_utils",
    "fms_fsdp/policies/ac_handler.py": "fms_fsdp.policies.ac_handler",
    "fms_fsdp/policies/mixed_precision.py": "fms_fsdp.policies.mixed_precision",
    "fms_fsdp/policies/wrapping.py": "fms_fsdp.policies.wrapping",
    "fms_fsdp/policies/param_init.py": "fms_fsdp.policies.param_init",
    "speculator/train_speculator_utils.py": "speculator.train_speculator_utils",
    "speculator/train_speculator.py": "speculator.train_speculator",
    "fms_to_hf_llama.py": "fms_to_hf_llama",
    "fms_to_hf_mamba.py": "fms_to_hf_mamba",
    "main_training_llama.py": "main_training_llama",
    "main_training_mamba.py": "main_training_mamba",
}


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("path,mod", sorted(MODULES.items()))
def test_public_names_of_reference_module_exist(path, mod):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    ours = importlib.import_module(mod)
    missing = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and not node.name.startswith("_"):
            if not hasattr(ours, node.name):
                missing.append(node.name)
            elif isinstance(node, ast.ClassDef):
                cls = getattr(ours, node.name)
                missing += [f"{node.name}.{b.name}" for b in node.body
                            if isinstance(b, ast.FunctionDef) and not b.name.startswith("_") and not hasattr(cls, b.name)]
        elif isinstance(node, ast.Assign):
            missing += [t.id for t in node.targets
                        if isinstance(t, ast.Name) and not t.id.startswith("_") and not hasattr(ours, t.id)]
    assert not missing, f"{mod} lacks {missing}"


def _positional_names(fn_node):
    a = fn_node.args
    return [x.arg for x in a.posonlyargs + a.args if x.arg not in ("self", "cls")], [x.arg for x in a.kwonlyargs]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("path,mod", sorted(MODULES.items()))
def test_signatures_accept_the_reference_call_forms(path, mod):
    """Every public function / method (and ``__init__``) takes the reference's parameters under the same names and, unless it
    takes ``*args``, at the same positions -- so both keyword and positional call sites of reference users keep working."""
    import inspect
    tree = ast.parse(open(os.path.join(REF, path)).read())
    ours = importlib.import_module(mod)
    problems = []

    def check(node, obj, label):
        try:
            params = list(inspect.signature(obj).parameters.values())
        except (TypeError, ValueError):
            return
        names = [p.name for p in params if p.name not in ("self", "cls")]
        var_kw = any(p.kind == p.VAR_KEYWORD for p in params)
        var_pos = any(p.kind == p.VAR_POSITIONAL for p in params)
        pos, kwonly = _positional_names(node)
        for i, n in enumerate(pos):
            if n not in names:
                if not var_kw:
                    problems.append(f"{label}: no parameter {n}")
            elif names.index(n) != i and not var_pos:
                problems.append(f"{label}: {n} is parameter {names.index(n)}, reference has it at {i}")
        problems.extend(f"{label}: no keyword {n}" for n in kwonly if n not in names and not var_kw)

    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and not node.name.startswith("_") and hasattr(ours, node.name):
            check(node, getattr(ours, node.name), node.name)
        elif isinstance(node, ast.ClassDef) and hasattr(ours, node.name):
            cls = getattr(ours, node.name)
            for b in node.body:
                if isinstance(b, ast.FunctionDef) and (b.name == "__init__" or not b.name.startswith("_")) and hasattr(cls, b.name):
                    check(b, getattr(cls, b.name), f"{node.name}.{b.name}")
    assert not problems, f"{mod}: {problems}"
