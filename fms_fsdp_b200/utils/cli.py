"""Tiny ``fire``-compatible argv parser (``fire`` is not in the image).

Semantics the entry points rely on:
  * ``--key=value`` and ``--key value`` -> kwargs; bare ``--flag`` -> True, ``--noflag`` -> False
  * values are literal-evaluated (``1`` -> int, ``3e-4`` -> float, ``True`` -> bool, ``None``),
    anything that is not a Python literal stays a string -- so ``--selective_checkpointing=1/3``
    arrives as the *string* "1/3" exactly as it did under fire.
"""
from __future__ import annotations

import ast
import sys
from typing import Any, Callable, Dict, List, Optional


def _literal(text: str) -> Any:
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        low = text.lower()
        if low == "true":
            return True
        if low == "false":
            return False
        if low in ("none", "null"):
            return None
        return text


def parse_argv(argv: Optional[List[str]] = None) -> Dict[str, Any]:
    argv = list(sys.argv[1:] if argv is None else argv)
    out: Dict[str, Any] = {}
    i = 0
    while i < len(argv):
        tok = argv[i]
        if not tok.startswith("-"):
            raise SystemExit(f"unexpected positional argument {tok!r}; use --key=value")
        key = tok.lstrip("-")
        if "=" in key:
            key, _, val = key.partition("=")
            out[key.replace("-", "_")] = _literal(val)
            i += 1
            continue
        key = key.replace("-", "_")
        if i + 1 < len(argv) and not argv[i + 1].startswith("--"):
            out[key] = _literal(argv[i + 1])
            i += 2
            continue
        if key.startswith("no") and len(key) > 2:
            out[key[2:]] = False
        else:
            out[key] = True
        i += 1
    return out


def run(main: Callable, argv: Optional[List[str]] = None):
    """Drop-in for ``fire.Fire(main)``."""
    return main(**parse_argv(argv))
