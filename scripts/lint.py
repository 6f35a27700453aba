"""Dependency-free static checks run by CI (`.github/workflows/lint.yml`) and runnable anywhere:
syntax, unused imports, names that are loaded but never bound anywhere in their module, and a line-length cap.

    python scripts/lint.py            # exit status 1 on findings
"""
import ast
import builtins
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = ("baseline/_ref", ".git", "gpurun_out", "__pycache__", "build")
MAX_LINE = 140
BUILTINS = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}


def files():
    for dirpath, _, names in os.walk(ROOT):
        rel = os.path.relpath(dirpath, ROOT)
        if any(rel == s or rel.startswith(s + os.sep) or (os.sep + s) in (os.sep + rel) for s in SKIP):
            continue
        for n in names:
            if n.endswith(".py"):
                yield os.path.join(dirpath, n)


def check(path):
    src = open(path).read()
    rel = os.path.relpath(path, ROOT)
    out = []
    try:
        tree = ast.parse(src)
    except SyntaxError as e:
        return [f"{rel}:{e.lineno}: syntax error: {e.msg}"]
    lines = src.splitlines()
    capped = not rel.startswith(("scripts" + os.sep, "tests" + os.sep))     # one-off diagnostics and tests may run long
    for i, line in enumerate(lines if capped else [], 1):
        if len(line) > MAX_LINE and "http" not in line and "noqa" not in line:
            out.append(f"{rel}:{i}: line longer than {MAX_LINE} characters ({len(line)})")
    bound, imported, loaded, exported = set(), {}, set(), set()
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(n.name)
        if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = n.args
            bound.update(x.arg for x in a.posonlyargs + a.args + a.kwonlyargs)
            bound.update(x.arg for x in (a.vararg, a.kwarg) if x)
        elif isinstance(n, ast.Name):
            (bound if isinstance(n.ctx, (ast.Store, ast.Del)) else loaded).add(n.id)
        elif isinstance(n, ast.Import):
            for a in n.names:
                imported[(a.asname or a.name).split(".")[0]] = n.lineno
        elif isinstance(n, ast.ImportFrom):
            for a in n.names:
                if a.name != "*":
                    imported[a.asname or a.name] = n.lineno
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
        elif isinstance(n, (ast.Global, ast.Nonlocal)):
            bound.update(n.names)
        elif isinstance(n, ast.Assign):
            for t in n.targets:
                if isinstance(t, ast.Name) and t.id == "__all__":
                    try:
                        exported = set(ast.literal_eval(n.value))
                    except Exception:
                        pass
    bound.update(imported)
    for name in sorted(loaded - bound - BUILTINS):
        out.append(f"{rel}: name {name!r} is used but never defined in the module")
    if not rel.endswith("__init__.py"):
        for name, ln in sorted(imported.items(), key=lambda kv: kv[1]):
            if name in loaded or name in exported or name == "annotations" or "noqa" in lines[ln - 1]:
                continue
            if src.count(name) > 1:          # referenced in a string annotation / docstring-driven registry
                continue
            out.append(f"{rel}:{ln}: unused import {name}")
    return out


if __name__ == "__main__":
    findings = [f for p in sorted(files()) for f in check(p)]
    print("\n".join(findings) if findings else "lint: clean")
    sys.exit(1 if findings else 0)
