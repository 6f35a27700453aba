"""Per-kernel SASS listings of the hot kernels (the ones bench.py launches), condensed: the full instruction stream of
each kernel's tensor-core / TMA / TMEM / peer-memory instructions IN ORDER with their addresses, plus a mnemonic
histogram -- evidence that the code is tcgen05 / TMA native (profiles/sass/*.txt).   python scripts/sass_listing.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fms_fsdp_b200", "_C.so")
OUT = os.path.join(ROOT, "profiles", "sass")
# (file tag, regex on the demangled kernel name)
HOT = [
    ("gemm2_nt_store", r"gemm2_bf16_tcgen05<false, false, 0, __nv_bfloat16, false>"),
    ("gemm2_nt_rope_ag", r"gemm2_bf16_tcgen05<false, false, 3, __nv_bfloat16, true>"),
    ("gemm2_nt_swiglu_ag", r"gemm2_bf16_tcgen05<false, false, 5, __nv_bfloat16, true>"),
    ("gemm2_nt_residual_ag", r"gemm2_bf16_tcgen05<false, false, 1, __nv_bfloat16, true>"),
    ("gemm2_nn_dgrad_ag", r"gemm2_bf16_tcgen05<false, true, 0, __nv_bfloat16, true>"),
    ("gemm2_tn_wgrad_push_ag", r"gemm2_bf16_tcgen05<true, true, 4, __nv_bfloat16, true>"),
    ("gemm2_tn_wgrad_push", r"gemm2_bf16_tcgen05<true, true, 4, __nv_bfloat16, false>"),
    ("gemm2_fp8_e4m3", r"gemm2_bf16_tcgen05<false, false, 7, __nv_bfloat16, false, true>"),
    ("quant_rowwise_e4m3", r"quant_rowwise_e4m3_kernel"),
    ("attn_fwd2", r"attn_fwd2_kernel<128, 0>"),
    ("attn_bwd3_dkdv", r"attn_bwd3_kernel<128, 0, 1>"),
    ("attn_bwd3_dq", r"attn_bwd3_kernel<128, 1, 1>"),
    ("attn_delta", r"attn_delta_kernel<128>"),
    ("reduce_scatter_bf16_w8", r"reduce_scatter_kernel<true, 8>"),
    ("signal_barrier", r"signal_barrier_kernel"),
    ("scalar_allreduce", r"scalar_allreduce_kernel"),
    ("p2p_allgather", r"p2p_allgather_kernel"),
    ("adamw_f32grad", r"adamw_kernel<float>"),
    ("rmsnorm_fwd", r"rmsnorm_fwd_kernel<2>"),
    ("rmsnorm_bwd", r"rmsnorm_bwd_kernel<2>"),
    ("swiglu_bwd", r"swiglu_bwd_kernel"),
    ("ce_grad_inplace", r"ce_grad_inplace_kernel"),
    ("embedding_bwd", r"embedding_bwd_kernel<int>"),
]
KEY = re.compile(r"\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|UTCBAR|LDTM|STTM|UTCATOMSWS|SYNCS|MUFU|FFMA2|FADD2|FMUL2|"
                 r"LDG\.E\.128\.STRONG\.SYS|ST\.E\.STRONG\.SYS|STG\.E\.STRONG\.SYS|LD\.E\.STRONG\.SYS|ATOM|RED|ELECT|FENCE|MEMBAR|ERRBAR)\S*")


def main():
    os.makedirs(OUT, exist_ok=True)
    names = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    # split per function
    blocks = re.split(r"\n\s*Function : ", names)
    demangle = lambda m: subprocess.run(["c++filt", m], capture_output=True, text=True).stdout.strip()
    index = []
    for blk in blocks[1:]:
        mangled = blk.split("\n", 1)[0].strip()
        index.append((mangled, demangle(mangled), blk))
    summary = []
    for tag, pat in HOT:
        hits = [(m, d, b) for m, d, b in index if re.search(pat.replace("<", r"<\(?(?:int\)|bool\))?").replace(", ", r",\s*\(?(?:int\)|bool\))?"), d) or pat in d.replace("(int)", "").replace("(bool)", "")]
        if not hits:
            summary.append(f"{tag}: NOT FOUND ({pat})")
            continue
        mangled, dem, blk = hits[0]
        insts = []
        for line in blk.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                insts.append((m.group(1), m.group(2).strip()))
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for _, i in insts)
        with open(os.path.join(OUT, tag + ".txt"), "w") as f:
            f.write(f"kernel   : {dem}\nmangled  : {mangled}\ninstructions: {len(insts)}\n\n")
            f.write("-- tensor-core / TMA / TMEM / barrier / peer-memory instructions, in program order --\n")
            for addr, ins in insts:
                if KEY.search(ins):
                    f.write(f"  /*{addr}*/  {ins}\n")
            f.write("\n-- mnemonic histogram --\n")
            for k, v in hist.most_common():
                f.write(f"  {v:6d}  {k}\n")
        keyc = collections.Counter(KEY.search(i).group(1) for _, i in insts if KEY.search(i))
        summary.append(f"{tag}: {len(insts)} instructions; " + ", ".join(f"{k} x{v}" for k, v in keyc.most_common(8)))
    with open(os.path.join(OUT, "INDEX.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
