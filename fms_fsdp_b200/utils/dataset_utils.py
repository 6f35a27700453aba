"""Stateful, rescalable, streaming dataset stack (written from the behavioural spec in SURVEY.md §2.3).

What is shared with the reference (``fms_fsdp/utils/dataset_utils.py``) is the *contract*, because loader checkpoints
must be interchangeable and its users construct these classes by name:

  * the 13 public classes and their constructor signatures;
  * the on-disk state: ``loader_state_<rank>.pth`` holding ``{"<ClassName>.<field>": value}`` with the reference's
    field names, ``<save>/checkpoints/step_<N>_ckp/`` directories, ``meta/*counts*.csv`` document counts;
  * the sampling arithmetic that decides WHICH document comes next (fragment ownership, ``random.seed(seed+rank)``
    file order, the a=5 / c=2(rank+seed)+1 LCG document order, ``torch.multinomial`` over documents remaining,
    largest-deficit corpus choice), so a resumed job continues the same stream.

How it is built is this repo's own: the ownership / resharding arithmetic lives in pure functions
(``owned_range``, ``covering_range``, ``take_owned``), persisted fields are declared with ``_persist`` instead of
hand-maintained lists, wrappers customise one hook (``_setup_children``) instead of re-implementing ``setup``, the
document reader is split into a precomputed ``_ReadPlan`` (bisect lookup), an ``_LcgOrder`` stepper and a chunk
geometry helper, and the packers are small explicit state machines.

Design rules kept from the reference's doc (``:19-34``): no communication between loader workers, pipelines are chains
of generators, state functions recurse through the chain, and rescalability comes from splitting state into scalars
(meaningful only for an unchanged world size) and lists (re-partitioned by fractional ownership).

Pipeline (inner -> outer; all off the GPU critical path, inside DataLoader workers):
``StreamingDocDataset -> ScalableShardDataset -> SamplingDataset -> BufferDataset -> PreloadBufferDataset ->
PreprocessDataset -> CheckpointDataset``.
"""
from __future__ import annotations

import bisect
import csv
import logging
import math
import os
import random
import time
from copy import deepcopy
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Set, Tuple, Union

import torch
import torch.utils.data as data

try:  # pyarrow is only needed by the file handlers
    import pyarrow as pa
    import pyarrow.parquet as pq
except Exception:  # pragma: no cover
    pa = pq = None


# =========================================================================== ownership arithmetic
def owned_range(n_items: int, rank: int, worldsize: int) -> Tuple[int, int]:
    """Half-open index range of the items rank ``rank`` owns outright when ``n_items`` are dealt contiguously."""
    return (rank * n_items) // worldsize, ((rank + 1) * n_items) // worldsize


def covering_range(n_items: int, rank: int, worldsize: int) -> Tuple[int, int]:
    """Half-open range of the items the rank owns *any fraction of* (used to pick which checkpoint shards to read)."""
    return (n_items * rank) // worldsize, -((-n_items * (rank + 1)) // worldsize)


def _shard_partition(itemlist: List[Any], rank: int, worldsize: int) -> List[Any]:
    """Reference: ``fms_fsdp/utils/dataset_utils.py:45-51``."""
    lo, hi = owned_range(len(itemlist), rank, worldsize)
    return itemlist[lo:hi]


def _shard_inclusive(itemlist: List[Any], rank: int, worldsize: int) -> List[Any]:
    """Reference: ``fms_fsdp/utils/dataset_utils.py:54-61``."""
    lo, hi = covering_range(len(itemlist), rank, worldsize)
    return itemlist[lo:hi]


def take_owned(covering: Sequence[Sequence[Any]], load_worldsize: int, rank: int, worldsize: int) -> List[Any]:
    """Rescale a list-valued state.  The checkpoint held ``load_worldsize`` equal-length lists; conceptually they are
    concatenated and re-dealt over ``worldsize`` ranks.  ``covering`` are only the checkpoint lists that overlap this
    rank (``_shard_inclusive`` of the full set); return the items this rank owns."""
    per = len(covering[0])
    if any(len(c) != per for c in covering):
        bad = next(i for i, c in enumerate(covering) if len(c) != per)
        raise AssertionError(f"Shard {bad} with length {len(covering[bad])} does not match expected {per}")
    first_shard = covering_range(load_worldsize, rank, worldsize)[0]
    lo, hi = owned_range(load_worldsize * per, rank, worldsize)
    base = first_shard * per
    return [covering[(g - base) // per][(g - base) % per] for g in range(lo, hi)]


def _latest_step_dir(path: str) -> str:
    from fms_fsdp_b200.utils.checkpointing_utils import get_latest
    return get_latest(path, key=lambda p: int(p.split("_")[-2]))


# ======================================================================================= base classes
class _StatefulDataset(data.IterableDataset):
    """Iterable dataset whose position can be saved, restored and re-partitioned.

    Subclasses call ``_persist(scalars=..., lists=...)``; the two groups are exposed as ``state_params`` /
    ``reshard_params`` (the reference's attribute names)."""

    def __init__(self, datapath: Optional[str], rank: int, worldsize: int):
        assert rank >= 0, f"Rank {rank} must be a positive integer"
        assert worldsize > rank, f"Worldsize {worldsize} must be greater than rank {rank}"
        assert datapath is None or (os.path.isdir(datapath) and len(os.listdir(datapath)) > 0), \
            f"Data path {datapath} must be a non-empty folder or None"
        self.state_params: List[str] = []
        self.reshard_params: List[str] = []
        self.datapath, self.rank, self.worldsize = datapath, rank, worldsize
        self.local_worldsize = -1          # -1: DataLoader workers not folded into (rank, worldsize) yet
        self.load_worldsize = worldsize    # world size of the checkpoint being loaded
        self.is_setup = False

    def _persist(self, scalars: Sequence[str] = (), lists: Sequence[str] = ()):
        self.state_params = list(scalars)
        self.reshard_params = list(lists)

    # ---- deferred initialisation: runs inside the DataLoader worker, where the worker id is known
    def _fold_in_workers(self):
        if self.local_worldsize != -1:
            return                          # an enclosing wrapper already did it and pushed the result down
        info = data.get_worker_info()
        n = 1 if info is None else info.num_workers
        self.local_worldsize = n
        if n > 1:                           # every DataLoader worker is a loader rank of its own
            self.rank = self.rank * n + info.id
            self.worldsize *= n

    def setup(self):
        if not self.is_setup:
            self.is_setup = True
            self._fold_in_workers()

    # ---- state
    def statename(self, x: str) -> str:
        """One key space per class: a layer type may appear once per pipeline."""
        return f"{type(self).__name__}.{x}"

    def _own_state(self) -> Dict[str, Any]:
        return {self.statename(k): getattr(self, k) for k in (*self.state_params, *self.reshard_params)}

    def state_dict(self) -> Dict[str, Any]:
        self.setup()
        return self._own_state()

    def _restore(self, covering: List[Dict[str, Any]]):
        if self.load_worldsize == self.worldsize:       # unchanged job size: everything comes back verbatim
            src = covering[0]
            for k in (*self.state_params, *self.reshard_params):
                setattr(self, k, src[self.statename(k)])
        else:                                           # rescaled: scalars are dropped, lists are re-dealt
            for k in self.reshard_params:
                name = self.statename(k)
                setattr(self, k, take_owned([sd[name] for sd in covering], self.load_worldsize, self.rank, self.worldsize))

    def load_state_dict(self, state_dicts: List[Dict[str, Any]], sharded_input: bool = False):
        """``state_dicts``: every checkpoint shard, or (``sharded_input``) only those overlapping this rank.  Returns
        the overlapping shards so wrappers can hand them down."""
        self.setup()
        if not sharded_input:
            self.load_worldsize = len(state_dicts)
            state_dicts = _shard_inclusive(state_dicts, self.rank, self.worldsize)
        self._restore(state_dicts)
        return state_dicts

    def load_from_path(self, path: str):
        """Read just the ``loader_state_<r>.pth`` files whose rank span overlaps ours."""
        assert os.path.exists(path), "Specified checkpoint does not exist"
        assert not os.path.isfile(path), "Checkpoint should be a folder of shard states"
        files = sorted((f for f in os.listdir(path) if "loader" in f), key=lambda f: int(f.split("_")[2][:-4]))
        assert len(files) > 0, "Checkpoint directory must contain checkpoint files with 'loader' in the name"
        self.load_worldsize = len(files)
        wanted = _shard_inclusive(files, self.rank, self.worldsize)
        self.load_state_dict([torch.load(os.path.join(path, f), weights_only=False) for f in wanted], True)

    def save_to_path(self, path: str):
        os.makedirs(path, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(path, f"loader_state_{self.rank}.pth"))


class _WrapperDataset(_StatefulDataset):
    """A layer around one sub-dataset.  ``_setup_children`` is the single customisation point: the default prepares the
    wrapped dataset itself; layers that fan out (logical shards, corpora) build their clones there instead.
    Reference: ``fms_fsdp/utils/dataset_utils.py:223-280``."""

    def __init__(self, dataset: _StatefulDataset):
        self.dataset = dataset
        super().__init__(dataset.datapath, dataset.rank, dataset.worldsize)

    def _place(self, child: _StatefulDataset, **overrides):
        """Give ``child`` this layer's (worker-adjusted) coordinates."""
        child.datapath, child.rank, child.worldsize, child.local_worldsize = (
            self.datapath, self.rank, self.worldsize, self.local_worldsize)
        for k, v in overrides.items():
            setattr(child, k, v)
        return child

    def _setup_children(self):
        self._place(self.dataset).setup()

    def setup(self):
        if not self.is_setup:
            super().setup()
            self._setup_children()

    def state_dict(self):
        self.setup()
        out = self.dataset.state_dict()
        own = self._own_state()
        for k in own.keys() & out.keys():
            logging.warning(f"Loader {self.rank}: flag {k} already present in state_dict, overwriting")
        out.update(own)
        return out

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        covering = super().load_state_dict(state_dicts, sharded_input)
        self.dataset.load_worldsize = self.load_worldsize
        self.dataset.load_state_dict(covering, True)
        return covering


# ======================================================================================= file handlers
class _ShardFileHandler:
    """Format adapter: which files qualify, how to open them, and how to count / fetch / slice documents.
    Reference: ``fms_fsdp/utils/dataset_utils.py:286-330``."""

    def is_legal(self, filepath: str) -> bool:
        return os.path.isfile(filepath)

    def open(self, path: str):
        raise NotImplementedError

    def length(self, path: str) -> int:
        raise NotImplementedError

    def get(self, reader, index: int, drop_tokens: Set):
        raise NotImplementedError

    def slice(self, doc, index: int, n_pull: int) -> List:
        raise NotImplementedError


def _has_ext(path: str, tag: str) -> bool:
    return tag in os.path.splitext(path)[1]


class ArrowHandler(_ShardFileHandler):
    """Pre-tokenised Arrow IPC shards, one RecordBatch per document with the tokens in ``col_name``.  Memory-mapped:
    only the slice of a document that a chunk needs is ever turned into Python objects.
    Reference: ``fms_fsdp/utils/dataset_utils.py:333-368``."""

    def __init__(self, col_name: str = "tokens"):
        self.col_name = col_name

    def is_legal(self, filepath: str) -> bool:
        return _has_ext(filepath, "arrow")

    def open(self, path: str):
        return pa.ipc.open_file(pa.memory_map(path))

    def length(self, path: str) -> int:
        return self.open(path).num_record_batches

    def get(self, reader, index: int, drop_tokens: Set):
        doc = reader.get_batch(index)[self.col_name]
        lo, hi = 0, len(doc)
        if hi > lo and doc[lo].as_py() in drop_tokens:
            lo += 1
        if hi > lo and doc[hi - 1].as_py() in drop_tokens:
            hi -= 1
        return doc.slice(lo, hi - lo)

    def slice(self, doc, index: int, n_pull: int) -> List:
        part = doc.slice(index, n_pull)
        try:
            # integer column without nulls: one vectorised conversion (15x faster than per-element ``as_py``)
            return part.to_numpy(zero_copy_only=True).tolist()
        except (pa.ArrowInvalid, pa.ArrowNotImplementedError, NotImplementedError, TypeError):
            return part.to_pylist()


class ParquetHandler(_ShardFileHandler):
    """HF-style parquet shards with a raw text column, tokenised on the fly.
    Reference: ``fms_fsdp/utils/dataset_utils.py:371-404``."""

    def __init__(self, tokenizer_path: str, col_name: str = "text"):
        from transformers import AutoTokenizer
        self.tokenizer = AutoTokenizer.from_pretrained(tokenizer_path)
        self.col_name = col_name

    def is_legal(self, filepath: str) -> bool:
        return _has_ext(filepath, "parquet")

    def open(self, path: str):
        return pq.read_pandas(path, columns=[self.col_name], partitioning=None)[self.col_name]

    def length(self, path: str) -> int:
        return pq.read_metadata(path).num_rows

    def get(self, reader, index: int, drop_tokens: Set):
        ids = self.tokenizer(str(reader[index]))["input_ids"]
        lo, hi = 0, len(ids)
        if hi > lo and ids[lo] in drop_tokens:
            lo += 1
        if hi > lo and ids[hi - 1] in drop_tokens:
            hi -= 1
        return ids[lo:hi]

    def slice(self, doc: List, index: int, n_pull: int) -> List:
        return doc[index: index + n_pull]


class AutoHandler(_ShardFileHandler):
    """Picks Arrow or parquet per file extension.
    Reference: ``fms_fsdp/utils/dataset_utils.py:407-457``."""

    def __init__(self, tokenizer_path: str, col_name: str = "text"):
        self.PHandler = ParquetHandler(tokenizer_path, col_name)
        self.AHandler = ArrowHandler()
        self.current: _ShardFileHandler = _ShardFileHandler()

    def _for(self, path: str) -> _ShardFileHandler:
        return self.AHandler if _has_ext(path, "arrow") else self.PHandler

    def is_legal(self, filepath: str) -> bool:
        return _has_ext(filepath, "parquet") or _has_ext(filepath, "arrow")

    def open(self, path: str):
        self.current = self._for(path)
        return self.current.open(path)

    def length(self, path: str) -> int:
        return self._for(path).length(path)

    def get(self, reader, index: int, drop_tokens: Set):
        return self.current.get(reader, index, drop_tokens)

    def slice(self, doc, index: int, n_pull: int) -> List:
        return self.current.slice(doc, index, n_pull)


# ===================================================================================== thin wrappers
class PreprocessDataset(_WrapperDataset):
    """Stateless map over the stream.
    Reference: ``fms_fsdp/utils/dataset_utils.py:463-488``."""

    def __init__(self, dataset: _StatefulDataset, aug_fn: Callable):
        super().__init__(dataset)
        self.aug_fn = aug_fn

    def __iter__(self):
        return map(self.aug_fn, iter(self.dataset))


class CheckpointDataset(_WrapperDataset):
    """Saves the loader state from INSIDE the worker every ``interval`` batches (``steps_per_batch`` items each) into
    ``<save>/checkpoints/step_<N>_ckp/`` -- the folder the model ``Checkpointer`` writes, so model and loader shards sit
    together -- and restores on setup: the newest checkpoint of the save folder if there is one (a restarted job), else
    the newest of the load folder with the step counter back at zero (someone else's checkpoint).
    Reference: ``fms_fsdp/utils/dataset_utils.py:491-618``."""

    def __init__(self, dataset: _StatefulDataset, load_path: str, interval: int, steps_per_batch: int = 1,
                 save_path: str = ""):
        super().__init__(dataset)
        self.interval, self.spb = interval, steps_per_batch
        self.load_path = os.path.join(load_path, "checkpoints")
        self.path = os.path.join(save_path, "checkpoints") if save_path else self.load_path
        self.step = 0        # batches emitted
        self.ministep = 0    # items of the current batch emitted

    def setup(self):
        if not self.is_setup:
            super().setup()
            self.load_from_path(self.load_path)

    def report(self, msg):
        if self.rank == 0:
            print(msg)

    def __iter__(self):
        self.setup()
        for item in self.dataset:
            yield item
            self.ministep += 1
            if self.ministep < self.spb:
                continue
            self.ministep = 0
            self.step += 1
            if self.step % self.interval == 0:
                self.save_to_path(os.path.join(self.path, f"step_{self.step}_ckp"))

    def _validate_ckp_path(self, path: str, verbose: bool = False) -> str:
        """Newest ``step_N_ckp`` folder under ``path`` that contains loader shards, '' if there is none.  Side effect:
        ``self.step = N``."""
        tell = self.report if verbose else (lambda m: None)
        if not os.path.exists(path) or not os.listdir(path):
            tell(f"  Dataset: No valid checkpoint detected at {path}, dataset starting from scratch.")
            return ""
        newest = _latest_step_dir(path)
        tell(f"Checkpoint detected at {newest}")
        if os.path.isfile(newest):
            tell(f"  Dataset: Detected checkpoint {newest} is a single file with no dataset info. Dataset starting from scratch.")
            return ""
        if not any("loader" in f for f in os.listdir(newest)):
            tell(f"  Dataset: Detected checkpoint {newest} exists but contains no dataset checkpoints. Dataset starting from scratch.")
            return ""
        self.step = int(newest.split("_")[-2])
        return newest

    def save_to_path(self, path: str):
        self.report(f"Saving dataset to {path}")
        t0 = time.time()
        super().save_to_path(path)
        self.report(f"Dataset successfully saved to {path}! Save time: {time.time() - t0}")

    def load_from_path(self, path: str):
        source = self._validate_ckp_path(self.path, False)
        if source:
            self.report(f"  Dataset: Detected a checkpoint in the save directory {source}. Restoring from this checkpoint.")
        else:
            source = self._validate_ckp_path(self.load_path, True)
            if not source:
                return
            self.step = 0     # keep the data position of the foreign checkpoint, count our own steps
        t0 = time.time()
        self.dataset.load_from_path(source)
        self.report(f"Dataset checkpoint loaded! Load time: {time.time() - t0}")


class PreloadBufferDataset(_WrapperDataset):
    """Local shuffle through one reservoir of ``window_size`` lines: while the reservoir is short it takes one extra
    line per draw; a uniformly random slot is emitted and refilled.  After a down-scale the (re-dealt) reservoir may be
    over-full: it then drains by one line per draw until it is back at ``window_size``.
    Reference: ``fms_fsdp/utils/dataset_utils.py:621-696``."""

    def __init__(self, dataset: _StatefulDataset, window_size: int):
        super().__init__(dataset)
        assert window_size > 1, f"Window size {window_size} must be greater than 1 for shuffling to occur"
        self.window_size = window_size
        self.generator = torch.Generator().manual_seed(self.rank)
        self.g_state = None
        self.buffer: List[Any] = []
        self._persist(scalars=["g_state"], lists=["buffer"])

    @property
    def buffer_size(self) -> int:
        return len(self.buffer)

    def __iter__(self):
        source = iter(self.dataset)
        pool = self.buffer
        while True:
            if len(pool) < self.window_size:
                pool.append(next(source))
            slot = int(torch.randint(len(pool), (1,), generator=self.generator))
            line = pool[slot]
            if len(pool) > self.window_size:
                last = pool.pop()
                if slot < len(pool):
                    pool[slot] = last
            else:
                pool[slot] = next(source)
            yield line
            pool = self.buffer      # a load_state_dict between two draws swaps the list

    def state_dict(self):
        self.g_state = self.generator.get_state()
        return super().state_dict()

    def load_state_dict(self, state_dicts, sharded_input=False):
        covering = super().load_state_dict(state_dicts, sharded_input)
        if self.g_state is not None:
            self.generator.set_state(self.g_state)
        self.buffer = list(self.buffer)
        return covering


class BufferDataset(_WrapperDataset):
    """Packs variable-length chunks into lines of exactly ``seq_len`` tokens.  ``pack_hard``: a chunk that does not fit
    is split across lines; otherwise the line is closed (EOS) and padded.  Optional per-line BOS / EOS are not
    duplicated when the token is already in place.  The carry-over is scalar state: it is dropped on a rescale.
    Reference: ``fms_fsdp/utils/dataset_utils.py:699-794``."""

    def __init__(self, dataset: _StatefulDataset, seq_len: int, pack_hard: bool, bos_token=None, eos_token=None,
                 pad_token=None):
        super().__init__(dataset)
        assert pack_hard or pad_token is not None, "Error: if using pads, you must supply a pad_token"
        self.len = seq_len
        self.pack_hard = pack_hard
        self.bos, self.eos, self.pad = bos_token, eos_token, pad_token
        self.buffer: List[Any] = []
        self._persist(scalars=["buffer"])

    def _close_line(self, tokens: List[Any]) -> Tuple[List[Any], List[Any]]:
        """First ``len`` tokens become the line; when an EOS has to overwrite its last token, that token moves to the
        front of the remainder."""
        line, rest = tokens[: self.len], tokens[self.len:]
        if self.eos is not None and line[-1] != self.eos:
            rest.insert(0, line[-1])
            line[-1] = self.eos
        return line, rest

    def _next_line(self, source: Iterator[List[Any]]) -> List[Any]:
        held, incoming = self.buffer, []
        while len(held) + len(incoming) < self.len:      # stop at the chunk that reaches (or overruns) the line end
            held = held + incoming
            incoming = next(source)
        if self.bos is not None and (not held or held[0] != self.bos):
            held = [self.bos] + held
        if len(held) >= self.len:                         # the carry-over alone fills the line
            line, rest = self._close_line(held)
            self.buffer = rest + incoming
        elif self.pack_hard:
            line, self.buffer = self._close_line(held + incoming)
        else:
            if self.eos is not None and held[-1] != self.eos:
                held = held + [self.eos]
            line = held + [self.pad] * (self.len - len(held)) if self.pad is not None else held
            self.buffer = incoming
        return line

    def __iter__(self):
        source = iter(self.dataset)
        while True:
            yield self._next_line(source)


# ===================================================================================== document reader
class _ReadPlan:
    """What one reader owns: ``(relative shard path, first doc, n docs)`` spans in visiting order, with running totals
    so that 'the i-th owned document' resolves by bisection."""

    def __init__(self, spans: List[Tuple[str, int, int]]):
        self.spans = spans
        self.ends: List[int] = []
        total = 0
        for _, _, n in spans:
            total += n
            self.ends.append(total)
        self.total = total

    def locate(self, i: int) -> Tuple[str, int, int]:
        assert i <= self.total, f"You have requested an illegal doc index {i}, docset length is {self.total}"
        k = bisect.bisect_right(self.ends, i)
        return self.spans[k]


class _LcgOrder:
    """Visit ``range(size)`` in the order of the full-period LCG x -> (5x + c) mod 2^ceil(log2 size), skipping values
    >= size.  c is odd, so the generator is a bijection: a shuffle that needs one integer of state."""

    def __init__(self, increment: int):
        self.c = increment

    def after(self, state: int, size: int) -> int:
        m = 1 << max(0, math.ceil(math.log2(size)))
        while True:
            state = (5 * state + self.c) % m
            if state < size:
                return state


class StreamingDocDataset(_StatefulDataset):
    """Distributed reader over a directory (tree) of shard files.

    Ownership: every shard file is cut into ``worldsize`` fragments; the ``n_files * worldsize`` fragments are dealt
    contiguously, so a rank holds one contiguous document range in each file it touches.  Order: the owned files are
    shuffled with ``random.seed(seed + rank)``; inside a file the documents follow an LCG permutation (``_LcgOrder``),
    so no index list is materialised.  A document is emitted as chunks of at most ``max_chunksize`` tokens; the first
    chunk carries the optional BOS, the last one the delimiter.  The reader resumes in the middle of a document and
    replays that document's already-emitted head chunks at the very end of the epoch.
    Reference: ``fms_fsdp/utils/dataset_utils.py:797-1145``."""

    def __init__(self, datapath: str, rank: int, worldsize: int, filehandler: _ShardFileHandler, delimiter_token: Any,
                 bos_token: Optional[Any] = None, strip_tokens: Optional[Set[Any]] = None, seed: int = 42,
                 min_length: int = 1, max_chunksize: int = 1024, verbose: bool = False):
        super().__init__(datapath, rank, worldsize)
        assert max_chunksize > 0, "Max chunksize must be a nonzero positive integer"
        self.filehandler, self.seed, self.verbose = filehandler, seed, verbose
        self.min_length, self.chunksize = min_length, max_chunksize
        self.eos, self.bos = delimiter_token, bos_token
        self.drop = set(strip_tokens) if strip_tokens else set()
        # position
        self.dataset = ""          # corpus name (last path component); checked on load
        self.docset_index = 0      # which owned document
        self.chunk_index = -1      # last chunk emitted of that document
        self.lcg_state = 0
        # statistics
        self.epochs_seen, self.tokens_seen, self.docs_seen, self.percent_seen = -1, 0, 0, 0
        self._persist(scalars=["dataset", "docset_index", "chunk_index", "epochs_seen", "tokens_seen", "docs_seen",
                               "percent_seen", "lcg_state"])
        self.docset: List[Tuple[str, int, int]] = []   # (shard, first doc, last doc) inclusive, visiting order
        self._plan = _ReadPlan([])
        self._len = 0
        self._order = _LcgOrder(1)

    # ---- planning
    def _legal_shards(self) -> List[str]:
        root = self.datapath
        rel = [os.path.relpath(os.path.join(d, f), root) for d, _, files in os.walk(root) for f in files
               if self.filehandler.is_legal(os.path.join(d, f))]
        return sorted(rel)      # the same order on every machine

    def _document_counts(self, corpus_parent: str, needed: Sequence[str]) -> Dict[str, int]:
        """Documents per shard: from ``<parent>/meta/*counts*.csv`` (columns dataset/filename, documents, tokens) when
        present, otherwise by opening the shards this reader needs."""
        meta = os.path.join(corpus_parent, "meta")
        tables = [f for f in os.listdir(meta) if "counts" in f and "csv" in f] if os.path.isdir(meta) else []
        if not tables:
            return {s: self.filehandler.length(os.path.join(self.datapath, s)) for s in needed}
        counts: Dict[str, int] = {}
        marker = "/" + self.dataset
        with open(os.path.join(meta, tables[0]), "r") as fh:
            for row in csv.DictReader(fh):
                name = row["dataset/filename"]
                at = name.find(marker)
                if at >= 0:
                    counts[name[at + len(marker) + 1:]] = int(row["documents"])
        return counts

    def _owned_spans(self, shards: List[str], counts: Dict[str, int]) -> List[Tuple[str, int, int]]:
        W, F = self.worldsize, len(shards)
        lo_frag, hi_frag = self.rank * F, (self.rank + 1) * F      # == owned_range(F * W, rank, W)
        spans = []
        for s in range(lo_frag // W, (hi_frag + W - 1) // W):
            f0 = max(lo_frag, s * W) - s * W
            f1 = min(hi_frag, (s + 1) * W) - s * W                 # exclusive
            n = counts[shards[s]]
            spans.append((shards[s], (n * f0) // W, (n * f1) // W - 1))
        return spans

    def setup(self):
        if self.is_setup:
            return
        super().setup()
        parent, name = os.path.split(self.datapath.rstrip(os.sep))
        self.dataset = name
        shards = self._legal_shards()
        W, F = self.worldsize, len(shards)
        touched = shards[(self.rank * F) // W: ((self.rank + 1) * F + W - 1) // W]
        self.docset = self._owned_spans(shards, self._document_counts(parent, touched))
        if self.verbose:
            logging.info(f"    Worker {self.rank} ingested {F} shard fragments from {self.dataset}")
        key = self.seed + self.rank
        random.seed(key)
        random.shuffle(self.docset)             # file order differs per worker ...
        self.lcg_state = key                    # ... and so does the document order inside the files
        self._order = _LcgOrder(2 * key + 1)
        self._plan = _ReadPlan([(s, lo, hi - lo + 1) for s, lo, hi in self.docset])
        self._len = self._plan.total

    # ---- chunk geometry
    def _n_chunks(self, doc) -> int:
        wrapped = len(doc) + (1 if self.bos is None else 2)       # + delimiter (+ BOS)
        return math.ceil(wrapped / self.chunksize) if wrapped >= self.min_length else 0

    def _chunk(self, doc, j: int, n_chunks: int) -> List[Any]:
        start, want = j * self.chunksize, self.chunksize
        if self.bos is not None:            # the BOS occupies slot 0 of chunk 0 and shifts everything after it
            start, want = (0, want - 1) if j == 0 else (start - 1, want)
        body = self.filehandler.slice(doc, start, want)
        self.tokens_seen += len(body)
        if j == 0 and self.bos is not None:
            body = [self.bos] + body
        if j == n_chunks - 1:
            body = body + [self.eos]
        return body

    # ---- iteration
    def __iter__(self):
        self.setup()
        resume_doc, resume_lcg = self.docset_index, self.lcg_state
        already_out = self.chunk_index + 1      # chunks of the resume document that were emitted before the save
        n_docs = self._len
        open_path, reader = None, None

        def fetch(doc_index: int):
            nonlocal open_path, reader
            shard, first, span = self._plan.locate(doc_index)
            path = os.path.join(self.datapath, shard)
            if path != open_path:
                if self.verbose:
                    logging.info(f"Worker {self.rank} opening new file {path}")
                reader, open_path = self.filehandler.open(path), path
            slot = self._order.after(self.lcg_state, span)
            return slot, self.filehandler.get(reader, first + slot, self.drop)

        if n_docs == 0:
            raise RuntimeError(f"{self.dataset}: reader {self.rank} of {self.worldsize} owns no documents and cannot be iterated "
                               f"(more readers than documents under {self.datapath}?)")
        while True:
            emitted_before = self.tokens_seen
            for step in range(n_docs):
                doc_index = (resume_doc + step) % n_docs
                if doc_index == 0:
                    self.epochs_seen += 1
                self.docset_index = doc_index
                slot, doc = fetch(doc_index)
                n_chunks = self._n_chunks(doc) if len(doc) else 0
                for j in range(already_out if step == 0 else 0, n_chunks):
                    self.chunk_index = j
                    if j == n_chunks - 1:
                        self.docs_seen += 1
                        self.percent_seen = self.docs_seen * 100 / (self._len + 1e-9)
                    yield self._chunk(doc, j, n_chunks)
                self.lcg_state = slot
            # close the lap: the head of the resume document that this lap skipped
            self.docset_index, self.lcg_state = resume_doc, resume_lcg
            if already_out:
                _, doc = fetch(resume_doc)
                n_chunks = self._n_chunks(doc) if len(doc) else 0
                for j in range(min(already_out, n_chunks)):
                    self.chunk_index = j
                    yield self._chunk(doc, j, n_chunks)
            if self.tokens_seen == emitted_before:
                # a whole lap over the owned documents produced nothing: the reference spins forever here
                raise RuntimeError(f"{self.dataset}: none of the {n_docs} documents of reader {self.rank} has at least "
                                   f"min_length={self.min_length} tokens (after stripping) -- nothing to read")

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        assert self.load_worldsize == self.worldsize, (
            f"StreamingDocDataset does not support rescaling (ckp size: {self.load_worldsize}, world size: "
            f"{self.worldsize}). Please use a ScalableShardDataset.")
        mine = self.dataset
        covering = super().load_state_dict(state_dicts, sharded_input)
        assert mine == self.dataset, f"Dataset mismatch: checkpoint contains {self.dataset}, expected {mine}"
        return covering


# ===================================================================================== rescalable layer
class ScalableShardDataset(_WrapperDataset):
    """Makes the reader rescalable: the corpus is cut into a FIXED number of logical readers (``n_logical_shards``,
    independent of the job size); a worker hosts ``n_logical_shards / worldsize`` of them and draws its next document
    from one picked with probability proportional to the documents it has left this epoch.  On a rescale the logical
    readers' states simply move to their new hosts.
    Reference: ``fms_fsdp/utils/dataset_utils.py:1148-1282``."""

    def __init__(self, dataset: StreamingDocDataset, delimiter_token: Any, n_logical_shards: int = 2048,
                 verbose: bool = False):
        super().__init__(dataset)
        assert n_logical_shards > 0, f"n_logical_shards {n_logical_shards} must be a positive integer"
        assert n_logical_shards % self.worldsize == 0, \
            f"World size {self.worldsize} must divide n_logical_shards {n_logical_shards} evenly"
        self.total_shards, self.delimiter, self.verbose = n_logical_shards, delimiter_token, verbose
        self.data: List[StreamingDocDataset] = []
        self.logicals_owned: List[int] = []
        self.n_logicals = 0
        self.generator: Optional[torch.Generator] = None
        self.n_docs_remaining: List[int] = []
        self.logical_shard_states = None
        self.current_reader = None      # logical reader with a document in flight
        self.g_state = None
        self._persist(scalars=["current_reader", "g_state"], lists=["n_docs_remaining", "logical_shard_states"])

    def _setup_children(self):
        lo, hi = owned_range(self.total_shards, self.rank, self.worldsize)
        self.logicals_owned = list(range(lo, hi))
        self.n_logicals = self.total_shards // self.worldsize
        assert len(self.logicals_owned) == self.n_logicals, "(world size * num workers) does not divide logical shards evenly"
        for k, logical in enumerate(self.logicals_owned):
            clone = self._place(deepcopy(self.dataset), rank=logical, worldsize=self.total_shards,
                                load_worldsize=self.total_shards, local_worldsize=1, verbose=self.rank == 0)
            clone.setup()
            self.data.append(clone)
            if self.verbose:
                logging.info(f"Worker {self.rank} assembled logical shard {logical}, {k + 1} of {self.n_logicals}")
        self.n_docs_remaining = [d._len for d in self.data]
        self.generator = torch.Generator().manual_seed(self.rank)

    def __iter__(self):
        self.setup()
        streams = [iter(d) for d in self.data]
        while True:
            pick = self.current_reader
            if pick is None:
                assert sum(self.n_docs_remaining) > 0, f"No documents detected in {self.datapath}"
                pick = int(torch.multinomial(torch.tensor(self.n_docs_remaining, dtype=torch.float), 1,
                                             generator=self.generator))
                self.current_reader = pick
            chunk = next(streams[pick])
            if chunk[-1] != self.delimiter:
                yield chunk             # more of this document follows: keep the reader selected
                continue
            self.current_reader = None
            self.n_docs_remaining[pick] -= 1
            if not any(self.n_docs_remaining):        # every logical reader finished its lap: new epoch
                self.n_docs_remaining = [d._len for d in self.data]
                self.generator.manual_seed(self.rank)
            yield chunk

    def state_dict(self):
        self.setup()
        self.g_state = self.generator.get_state()
        self.logical_shard_states = [d.state_dict() for d in self.data]
        return self._own_state()

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        covering = _StatefulDataset.load_state_dict(self, state_dicts, sharded_input)
        if self.g_state is not None:
            self.generator.set_state(self.g_state)
        for reader, saved in zip(self.data, self.logical_shard_states):
            reader.load_state_dict([saved], True)
        return covering


# ======================================================================================== corpus mixing
class SamplingDataset(_WrapperDataset):
    """Mixes corpora (sub-directories of ``datapath``) by token share: whenever a document ends, continue with the
    corpus whose share of the tokens emitted so far is furthest below its target weight.
    Reference: ``fms_fsdp/utils/dataset_utils.py:1285-1417``."""

    def __init__(self, datapath: str, dataset: Union[ScalableShardDataset, StreamingDocDataset], delimiter_token: Any,
                 datasets=None, weights=None, verbose=False):
        super().__init__(dataset)
        self.datapath, self.delimiter, self.verbose = datapath, delimiter_token, verbose
        if datasets is None:
            datasets = [d for d in os.listdir(datapath) if os.path.isdir(os.path.join(datapath, d)) and "meta" not in d]
        assert len(datasets) > 0, "You must specify at least one dataset"
        self.datasets = datasets
        if weights is None:
            weights = [1] * len(datasets)
        assert len(weights) == len(datasets), \
            f"Number of oversample weights {len(weights)} must match number of datasets {len(datasets)}"
        for w in weights:
            assert w > 0, f"Sampling rate {w} must be positive"
        self.weights = [w / sum(weights) for w in weights]
        self.data: List[_StatefulDataset] = []
        self.tokens_seen = [0] * len(datasets)
        self.current_iterator = -1      # corpus with a document in flight
        self._persist(scalars=["tokens_seen", "current_iterator"])

    def _setup_children(self):
        self.data = []
        for k, name in enumerate(self.datasets):
            sub = self._place(deepcopy(self.dataset), datapath=os.path.join(self.datapath, name))
            sub.setup()
            self.data.append(sub)
            if self.verbose:
                logging.info(f"Worker {self.rank} assembled subdataset iterator for {name}, {k + 1} of {len(self.datasets)}")

    def _most_underserved(self) -> int:
        total = sum(self.tokens_seen) + 1e-9
        best, best_gap = 0, -math.inf
        for i, (w, seen) in enumerate(zip(self.weights, self.tokens_seen)):
            gap = w - seen / total
            if gap >= best_gap:         # ties go to the later corpus
                best, best_gap = i, gap
        return best

    def __iter__(self):
        self.setup()
        streams = [iter(d) for d in self.data]
        while True:
            if self.current_iterator == -1:
                self.current_iterator = self._most_underserved()
            k = self.current_iterator
            chunk = next(streams[k])
            self.tokens_seen[k] += len(chunk)
            if chunk[-1] == self.delimiter:
                self.current_iterator = -1
            yield chunk

    def state_dict(self):
        self.setup()
        out = {self.statename("sample_iterator_states"): [d.state_dict() for d in self.data]}
        out.update(self._own_state())
        return out

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        covering = _StatefulDataset.load_state_dict(self, state_dicts, sharded_input)
        key = self.statename("sample_iterator_states")
        for i, sub in enumerate(self.data):
            sub.load_worldsize = self.load_worldsize
            sub.load_state_dict([sd[key][i] for sd in covering], True)
        return covering
