"""The shipped extension is Blackwell-native: every tensor-core instruction in it is a tcgen05 MMA (SASS ``UTCHMMA`` /
``UTCQMMA``), operands arrive by TMA (``UTMALDG``), accumulators are read from TMEM (``LDTM``), the pushing wgrad GEMM stores
with bulk copies (``UBLKCP``) -- and there is not a single legacy warp-level ``HMMA`` (``mma.sync`` / ``wmma``) in the binary.
Runs wherever ``cuobjdump`` and the built ``_C.so`` exist (no GPU needed)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fms_fsdp_b200", "_C.so")


@pytest.fixture(scope="module")
def kernels():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(SO) or not os.path.exists(exe):
        pytest.skip("needs the built extension and cuobjdump")
    out = subprocess.run([exe, "-sass", SO], capture_output=True, text=True, timeout=600).stdout
    assert "arch = sm_100a" in out
    per = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = []
        elif name and "/*" in line:
            per[name].append(line)
    return {k: "\n".join(v) for k, v in per.items()}


def _has(kernels, name_part, *mnemonics):
    hits = [k for k in kernels if name_part in k]
    assert hits, name_part
    for k in hits:
        for mn in mnemonics:
            assert re.search(mn, kernels[k]), (k, mn)
    return hits


def test_no_legacy_warp_level_mma_anywhere(kernels):
    legacy = [k for k, body in kernels.items() if re.search(r"(?<![A-Z])HMMA\.|(?<![A-Z])IMMA\.|(?<![A-Z])QMMA\.", body)]
    assert not legacy, legacy[:5]
    assert sum("UTCHMMA" in b or "UTCQMMA" in b for b in kernels.values()) >= 40     # GEMM + attention instantiations


def test_hot_kernels_use_tcgen05_tma_tmem(kernels):
    gemm = _has(kernels, "gemm2_bf16_tcgen05", r"UTCHMMA\.2CTA|UTCQMMA\.2CTA", r"UTMALDG\.2D", r"LDTM")
    assert len(gemm) >= 20
    _has(kernels, "attn_fwd2_kernel", r"UTCHMMA", r"UTMALDG", r"LDTM", r"MUFU\.EX2")
    _has(kernels, "attn_bwd3_kernel", r"UTCHMMA", r"UTMALDG", r"LDTM", r"STTM")
    # fp8 instantiation: kind::f8f6f4 on the CTA pair
    assert any("UTCQMMA.2CTA" in b for k, b in kernels.items() if "gemm2_bf16_tcgen05" in k)


def test_push_epilogue_and_peer_collectives(kernels):
    push = [k for k, b in kernels.items() if "gemm2_bf16_tcgen05" in k and "UBLKCP.G.S" in b]
    assert len(push) >= 2                                   # wgrad push, with and without the fused all-gather
    # comm warps of the all-gather GEMMs read peers with system-scope 16-byte loads
    assert any("LDG.E.128.STRONG.SYS" in b for k, b in kernels.items() if "gemm2_bf16_tcgen05" in k)
    _has(kernels, "signal_barrier_kernel", r"ST\S*\.STRONG\.SYS|STG\S*\.SYS", r"LD\S*\.STRONG\.SYS")
    _has(kernels, "reduce_scatter", r"LDG\.E\.128")
