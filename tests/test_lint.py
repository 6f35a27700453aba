"""The tree passes its own static checks (``scripts/lint.py``: syntax, unused imports, undefined names, line cap) -- the same
command the CI lint job runs."""
import os
import subprocess
import sys


def test_static_checks_are_clean():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "lint.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
