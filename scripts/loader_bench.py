"""CPU throughput of the data pipeline (tokens/s of ONE loader process, no worker pool): this repo's dataset stack vs the
unmodified reference's (baseline/_ref), same synthetic arrow corpus, same `get_data_loader` configuration.

    python scripts/loader_bench.py [--docs 20000] [--doc_len 1500] [--batches 300]

A training rank at the headline speed consumes ~25 k tokens/s; the loader must stay far above that per process."""
import argparse
import importlib
import os
import sys
import tempfile
import time

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_corpus(root, n_docs, doc_len, n_files=4, seed=0):
    rng = np.random.default_rng(seed)
    schema = pa.schema([pa.field("tokens", pa.uint32())])
    rows = []
    for ds in ("dataset_a", "dataset_b"):
        for f in range(n_files):
            path = os.path.join(root, ds, f"shard_{f}.arrow")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            ntok = 0
            with pa.ipc.new_file(path, schema) as w:
                for _ in range(n_docs // (2 * n_files)):
                    n = int(rng.integers(doc_len // 2, doc_len * 3 // 2))
                    w.write(pa.record_batch([pa.array(rng.integers(3, 32000, n, dtype=np.uint32))], schema=schema))
                    ntok += n
            rows.append(f"/{ds}/shard_{f}.arrow,{n_docs // (2 * n_files)},{ntok}")
    os.makedirs(os.path.join(root, "meta"))
    with open(os.path.join(root, "meta", "combined_counts.csv"), "w") as f:
        f.write("dataset/filename,documents,tokens\n" + "\n".join(rows) + "\n")


def run(which, data, batches, seq_len, batch_size):
    """The stack ``get_data_loader`` builds, assembled from the implementation's own public classes.  (The reference's
    ``get_data_loader`` cannot be called for arrow files at its HEAD: it hands the ArrowHandler CLASS, not an instance, to
    StreamingDocDataset -- ``is_legal() missing 1 required positional argument`` -- so both arms are assembled here.)"""
    import torch
    for m in [k for k in sys.modules if k == "fms_fsdp" or k.startswith("fms_fsdp.")]:
        del sys.modules[m]
    if which == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        D = importlib.import_module("fms_fsdp.utils.dataset_utils")
        dl = importlib.import_module("fms_fsdp.utils.dataloader_utils")
    else:
        sys.path.insert(0, ROOT)
        D = importlib.import_module("fms_fsdp_b200.utils.dataset_utils")
        dl = importlib.import_module("fms_fsdp_b200.utils.dataloader_utils")
    t0 = time.perf_counter()
    d = D.StreamingDocDataset(data, 0, 1, D.ArrowHandler(), 0, bos_token=None, strip_tokens={0}, min_length=3, seed=42)
    d = D.ScalableShardDataset(d, 0, n_logical_shards=64)
    d = D.SamplingDataset(data, d, 0, datasets=["dataset_a", "dataset_b"], weights=[2, 1], verbose=False)
    d = D.BufferDataset(d, seq_len + 1, bos_token=None, eos_token=None, pack_hard=True)
    d = D.PreloadBufferDataset(d, 10000)
    d = D.PreprocessDataset(d, torch.IntTensor)
    d = D.PreprocessDataset(d, dl.causal_lm)
    d = D.CheckpointDataset(d, tempfile.mkdtemp(), 10 ** 9, batch_size, tempfile.mkdtemp())
    loader = iter(torch.utils.data.DataLoader(d, num_workers=0, batch_size=batch_size))
    first = next(loader)
    t1 = time.perf_counter()
    for _ in range(batches):
        b = next(loader)
    t2 = time.perf_counter()
    sys.path.pop(0)
    return dict(setup_plus_first_batch_s=round(t1 - t0, 2), tokens_per_s=round(batches * batch_size * seq_len / (t2 - t1)),
                shape=tuple(b[0].shape), first_tokens=first[0][0, :6].tolist())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=20000)
    ap.add_argument("--doc_len", type=int, default=1500)
    ap.add_argument("--batches", type=int, default=300)
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    data = tempfile.mkdtemp(dir=os.path.join(ROOT, "gpurun_out"))
    make_corpus(data, a.docs, a.doc_len)
    out = {}
    for which in ("ours", "reference"):
        try:
            out[which] = run(which, data, a.batches, a.seq, a.batch)
        except Exception as e:   # the reference arm needs baseline/_ref
            out[which] = {"unavailable": repr(e)[:200]}
        print(which, out[which], flush=True)
