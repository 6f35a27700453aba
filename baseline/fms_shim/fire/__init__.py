"""Stand-in for the `fire` CLI package (absent offline): `Fire(fn)` -> fn(**kwargs from --k=v argv)."""
import ast
import sys


def Fire(fn):
    kw = {}
    for a in sys.argv[1:]:
        if a.startswith("--") and "=" in a:
            k, v = a[2:].split("=", 1)
            try:
                v = ast.literal_eval(v)
            except Exception:
                pass
            kw[k] = v
    return fn(**kw)
