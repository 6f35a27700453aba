"""Stateful, rescalable, streaming dataset stack.

Capability parity with reference ``fms_fsdp/utils/dataset_utils.py``: the same
13 public classes, the same pipeline semantics and the same on-disk loader-state format
(``loader_state_<rank>.pth`` holding ``{"ClassName.flag": value}``), so loader checkpoints written by
either implementation resume in the other.  Design rules (reference ``:19-34``):

  1. no communication between loader workers -- every worker derives its partition from
     (rank, worldsize) alone;
  2. a pipeline is a chain of wrappers around one base reader, each a python generator;
  3. ``state_dict`` / ``load_state_dict`` recurse through the chain;
  4. rescalability: ``state_params`` (scalars; only meaningful for an unchanged world size) versus
     ``reshard_params`` (lists; re-partitioned across a new world size by fractional ownership).

Pipeline layers (inner -> outer), all plain-Python and off the GPU critical path (they run in
DataLoader worker processes and feed pinned-memory batches to the training loop):
``StreamingDocDataset -> ScalableShardDataset -> SamplingDataset -> BufferDataset ->
PreloadBufferDataset -> PreprocessDataset -> CheckpointDataset``.
"""
from __future__ import annotations

import csv
import logging
import math
import os
import random
import time
from copy import deepcopy
from typing import Any, Callable, Dict, List, Optional, Sequence, Set, Tuple, Union

import torch
import torch.utils.data as data

try:  # pyarrow / transformers are only needed by the file handlers
    import pyarrow as pa
    import pyarrow.parquet as pq
except Exception:  # pragma: no cover
    pa = pq = None


# ------------------------------------------------------------------------------------- partitioning
def _shard_partition(itemlist: List[Any], rank: int, worldsize: int) -> List[Any]:
    """Contiguous integer partition: rank r owns [r*n//w, (r+1)*n//w)."""
    n = len(itemlist)
    return itemlist[(rank * n) // worldsize: ((rank + 1) * n) // worldsize]


def _shard_inclusive(itemlist: List[Any], rank: int, worldsize: int) -> List[Any]:
    """Span of items this rank owns *at least a fraction of* (floor/ceil of the fractional bounds)."""
    n = len(itemlist)
    return itemlist[math.floor(n * rank / worldsize): math.ceil(n * (rank + 1) / worldsize)]


def _latest_step_dir(path: str) -> str:
    from fms_fsdp_b200.utils.checkpointing_utils import get_latest
    return get_latest(path, key=lambda p: int(p.split("_")[-2]))


# ------------------------------------------------------------------------------------------- bases
class _StatefulDataset(data.IterableDataset):
    """Iterable dataset with (recursive, reshardable) state.

    Subclasses list the attribute names to persist in ``state_params`` (dropped when the world size
    changes) and ``reshard_params`` (lists that are re-split over the new world size).
    """

    def __init__(self, datapath: Optional[str], rank: int, worldsize: int):
        assert rank >= 0, f"Rank {rank} must be a positive integer"
        assert worldsize > rank, f"Worldsize {worldsize} must be greater than rank {rank}"
        assert datapath is None or (os.path.isdir(datapath) and len(os.listdir(datapath)) > 0), \
            f"Data path {datapath} must be a non-empty folder or None"
        self.state_params: List[str] = []
        self.reshard_params: List[str] = []
        self.datapath = datapath
        self.rank = rank
        self.worldsize = worldsize
        self.local_worldsize = -1
        self.load_worldsize = worldsize
        self.is_setup = False

    # -- deferred, rank-dependent initialisation (runs inside the DataLoader worker)
    def setup(self):
        if self.is_setup:
            return
        self.is_setup = True
        if self.local_worldsize == -1:  # not yet folded in by an enclosing wrapper
            info = data.get_worker_info()
            if info is None or info.num_workers == 1:
                self.local_worldsize = 1
            else:
                # every DataLoader worker becomes its own loader rank
                self.local_worldsize = info.num_workers
                self.worldsize *= self.local_worldsize
                self.rank = self.local_worldsize * self.rank + info.id

    def statename(self, x: str) -> str:
        # one key-space per class => a layer type may appear once per pipeline
        return self.__class__.__name__ + "." + x

    def state_dict(self) -> Dict[str, Any]:
        self.setup()
        return {self.statename(k): getattr(self, k) for k in self.state_params + self.reshard_params}

    def _reshard(self, sharded_list: List[List[Any]]) -> List[Any]:
        """``sharded_list`` = the checkpoint shards that overlap this rank (``_shard_inclusive`` of the
        global shard list, all of equal length).  Return exactly the flattened items this rank owns."""
        shard_len = len(sharded_list[0])
        for i, s in enumerate(sharded_list):
            assert len(s) == shard_len, f"Shard {i} with length {len(s)} does not match expected {shard_len}"
        dropped_shards = math.floor(self.load_worldsize * self.rank / self.worldsize)
        item_offset = shard_len * dropped_shards
        n_items = self.load_worldsize * shard_len
        lo = int(n_items * self.rank / self.worldsize) - item_offset
        hi = int(n_items * (self.rank + 1) / self.worldsize) - item_offset
        return [sharded_list[i // shard_len][i % shard_len] for i in range(lo, hi)]

    def load_state_dict(self, state_dicts: List[Dict[str, Any]], sharded_input: bool = False):
        """``state_dicts``: all checkpoint shards (``sharded_input=False``) or only those overlapping
        this rank.  Same world size -> restore everything from the single matching shard; different ->
        only ``reshard_params`` survive, re-partitioned."""
        self.setup()
        if not sharded_input:
            self.load_worldsize = len(state_dicts)
            state_dicts = _shard_inclusive(state_dicts, self.rank, self.worldsize)
        if self.load_worldsize == self.worldsize:
            src = state_dicts[0]
            for k in self.state_params + self.reshard_params:
                setattr(self, k, src[self.statename(k)])
        else:
            for k in self.reshard_params:
                setattr(self, k, self._reshard([sd[self.statename(k)] for sd in state_dicts]))
        return state_dicts

    def load_from_path(self, path: str):
        """Read only the ``loader_state_*.pth`` files whose rank span overlaps ours."""
        assert os.path.exists(path), "Specified checkpoint does not exist"
        assert not os.path.isfile(path), "Checkpoint should be a folder of shard states"
        files = [f for f in os.listdir(path) if "loader" in f]
        files.sort(key=lambda f: int(f.split("_")[2][:-4]))
        assert len(files) > 0, "Checkpoint directory must contain checkpoint files with 'loader' in the name"
        self.load_worldsize = len(files)
        mine = _shard_inclusive(files, self.rank, self.worldsize)
        states = [torch.load(os.path.join(path, f), weights_only=False) for f in mine]
        self.load_state_dict(states, True)

    def save_to_path(self, path: str):
        os.makedirs(path, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(path, f"loader_state_{self.rank}.pth"))


class _WrapperDataset(_StatefulDataset):
    """A layer around exactly one sub-dataset; state functions recurse into it."""

    def __init__(self, dataset: _StatefulDataset):
        self.dataset = dataset
        super().__init__(dataset.datapath, dataset.rank, dataset.worldsize)

    def setup(self):
        if self.is_setup:
            return
        super().setup()
        # rank/worldsize percolated up at construction; push the (worker-adjusted) values back down
        sub = self.dataset
        sub.datapath, sub.rank, sub.worldsize, sub.local_worldsize = (
            self.datapath, self.rank, self.worldsize, self.local_worldsize)
        sub.setup()

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        mine = super().load_state_dict(state_dicts, sharded_input)
        self.dataset.load_worldsize = self.load_worldsize
        self.dataset.load_state_dict(mine, True)
        return mine

    def state_dict(self):
        self.setup()
        out = self.dataset.state_dict()
        own = super().state_dict()
        for k in own:
            if k in out:
                logging.warning(f"Loader {self.rank}: flag {k} already present in state_dict, overwriting")
        out.update(own)
        return out


# ------------------------------------------------------------------------------------ file handlers
class _ShardFileHandler:
    """Format adapter: which files qualify, how to open them, count / fetch / slice documents."""

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


def _ext(path: str) -> str:
    return os.path.splitext(path)[1]


class ArrowHandler(_ShardFileHandler):
    """Pre-tokenised Arrow IPC shards: one RecordBatch per document, token list in ``col_name``.
    Memory-mapped, so only the requested chunk of a document is ever materialised."""

    def __init__(self, col_name: str = "tokens"):
        self.col_name = col_name

    def is_legal(self, filepath: str) -> bool:
        return "arrow" in _ext(filepath)

    def open(self, path: str):
        return pa.ipc.open_file(pa.memory_map(path))

    def length(self, path: str) -> int:
        return self.open(path).num_record_batches

    def get(self, reader, index: int, drop_tokens: Set):
        doc = reader.get_batch(index)[self.col_name]
        if len(doc) > 0 and doc[0].as_py() in drop_tokens:
            doc = doc.slice(1, len(doc) - 1)
        if len(doc) > 0 and doc[-1].as_py() in drop_tokens:
            doc = doc.slice(0, len(doc) - 1)
        return doc

    def slice(self, doc, index: int, n_pull: int) -> List:
        return doc.slice(index, n_pull).to_pylist()


class ParquetHandler(_ShardFileHandler):
    """HF-style parquet shards with a raw text column, tokenised on the fly."""

    def __init__(self, tokenizer_path: str, col_name: str = "text"):
        from transformers import AutoTokenizer
        self.tokenizer = AutoTokenizer.from_pretrained(tokenizer_path)
        self.col_name = col_name

    def is_legal(self, filepath: str) -> bool:
        return "parquet" in _ext(filepath)

    def open(self, path: str):
        return pq.read_pandas(path, columns=[self.col_name], partitioning=None)[self.col_name]

    def length(self, path: str) -> int:
        return pq.read_metadata(path).num_rows

    def get(self, reader, index: int, drop_tokens: Set):
        doc = self.tokenizer(str(reader[index]))["input_ids"]
        if len(doc) > 0 and doc[0] in drop_tokens:
            doc = doc[1:]
        if len(doc) > 0 and doc[-1] in drop_tokens:
            doc = doc[:-1]
        return doc

    def slice(self, doc: List, index: int, n_pull: int) -> List:
        return doc[index: index + n_pull]


class AutoHandler(_ShardFileHandler):
    """Dispatch per file extension between Arrow and parquet."""

    def __init__(self, tokenizer_path: str, col_name: str = "text"):
        self.PHandler = ParquetHandler(tokenizer_path, col_name)
        self.AHandler = ArrowHandler()
        self.current: _ShardFileHandler = _ShardFileHandler()

    def _pick(self, path: str) -> _ShardFileHandler:
        return self.AHandler if "arrow" in _ext(path) else self.PHandler

    def is_legal(self, filepath: str) -> bool:
        return "parquet" in _ext(filepath) or "arrow" in _ext(filepath)

    def open(self, path: str):
        self.current = self._pick(path)
        return self.current.open(path)

    def length(self, path: str) -> int:
        return self._pick(path).length(path)

    def get(self, reader, index: int, drop_tokens: Set):
        return self.current.get(reader, index, drop_tokens)

    def slice(self, doc, index: int, n_pull: int) -> List:
        return self.current.slice(doc, index, n_pull)


# ---------------------------------------------------------------------------------- simple wrappers
class PreprocessDataset(_WrapperDataset):
    """Stateless map of ``aug_fn`` over the stream."""

    def __init__(self, dataset: _StatefulDataset, aug_fn: Callable):
        super().__init__(dataset)
        self.aug_fn = aug_fn

    def __iter__(self):
        for item in iter(self.dataset):
            yield self.aug_fn(item)


class CheckpointDataset(_WrapperDataset):
    """Auto-save the loader state from inside the worker every ``interval`` *batches*
    (``steps_per_batch`` items each) into ``<save>/checkpoints/step_<N>_ckp/`` -- the directory the
    model ``Checkpointer`` uses, so model and loader shards co-locate -- and auto-load on setup:
    newest checkpoint in the save dir (job restart) else the load dir with the step count reset."""

    def __init__(self, dataset: _StatefulDataset, load_path: str, interval: int, steps_per_batch: int = 1,
                 save_path: str = ""):
        super().__init__(dataset)
        self.interval = interval
        self.spb = steps_per_batch
        load_path = os.path.join(load_path, "checkpoints")
        self.load_path = load_path
        self.path = os.path.join(save_path, "checkpoints") if len(save_path) > 0 else load_path
        self.step = 0
        self.ministep = 0

    def setup(self):
        if not self.is_setup:
            super().setup()
            self.load_from_path(self.load_path)

    def __iter__(self):
        self.setup()
        for item in iter(self.dataset):
            yield item
            self.ministep += 1
            if self.ministep == self.spb:
                self.ministep = 0
                self.step += 1
                if self.step % self.interval == 0:
                    self.save_to_path(os.path.join(self.path, f"step_{self.step}_ckp"))

    def report(self, msg):
        if self.rank == 0:
            print(msg)

    def _validate_ckp_path(self, path: str, verbose: bool = False) -> str:
        """Newest ``step_N_ckp`` folder under ``path`` that holds loader shards ('' if none);
        side effect: sets ``self.step`` to N."""
        def say(m):
            if verbose:
                self.report(m)
        if not os.path.exists(path) or len(os.listdir(path)) == 0:
            say(f"  Dataset: No valid checkpoint detected at {path}, dataset starting from scratch.")
            return ""
        latest = _latest_step_dir(path)
        say(f"Checkpoint detected at {latest}")
        if os.path.isfile(latest):
            say(f"  Dataset: Detected checkpoint {latest} is a single file with no dataset info."
                " Dataset starting from scratch.")
            return ""
        if not any("loader" in f for f in os.listdir(latest)):
            say(f"  Dataset: Detected checkpoint {latest} exists but contains no dataset checkpoints."
                " Dataset starting from scratch.")
            return ""
        self.step = int(latest.split("_")[-2])
        return latest

    def save_to_path(self, path: str):
        self.report(f"Saving dataset to {path}")
        t0 = time.time()
        super().save_to_path(path)
        self.report(f"Dataset successfully saved to {path}! Save time: {time.time() - t0}")

    def load_from_path(self, path: str):
        resume = self._validate_ckp_path(self.path, False)
        if resume:
            self.report(f"  Dataset: Detected a checkpoint in the save directory {resume}. Restoring from this checkpoint.")
            path = resume
        else:
            external = self._validate_ckp_path(self.load_path, True)
            if not external:
                return
            path = external
            self.step = 0  # someone else's checkpoint: keep the data position, restart the step count
        t0 = time.time()
        self.dataset.load_from_path(path)
        self.report(f"Dataset checkpoint loaded! Load time: {time.time() - t0}")


class PreloadBufferDataset(_WrapperDataset):
    """Local shuffle through one in/out buffer of ``window_size`` lines: grows two-at-a-time until
    full, then emits a uniformly random slot and refills it.  After a rescale the (resharded)
    buffer shrinks or re-grows back to ``window_size``."""

    def __init__(self, dataset: _StatefulDataset, window_size: int):
        super().__init__(dataset)
        assert window_size > 1, f"Window size {window_size} must be greater than 1 for shuffling to occur"
        self.window_size = window_size
        self.g_state = None
        self.generator = torch.Generator().manual_seed(self.rank)
        self.buffer: List[List[Any]] = []
        self.buffer_size = 0
        self.state_params = ["g_state"]
        self.reshard_params = ["buffer"]

    def _pad_buffer(self):
        if self.buffer_size < self.window_size:
            self.buffer += [[]] * (self.window_size - self.buffer_size)

    def __iter__(self):
        src = iter(self.dataset)
        while True:
            self._pad_buffer()
            if self.buffer_size < self.window_size:
                self.buffer[self.buffer_size] = next(src)
                self.buffer_size += 1
            i = torch.randint(self.buffer_size, (1,), generator=self.generator).item()
            out = self.buffer[i]
            if self.buffer_size > self.window_size:
                # oversized after a down-scale: drain instead of refilling
                self.buffer[i] = self.buffer[self.buffer_size - 1]
                self.buffer_size -= 1
            else:
                self.buffer[i] = next(src)
            yield out

    def state_dict(self):
        self.g_state = self.generator.get_state()
        self.buffer = self.buffer[: self.buffer_size]  # drop padding so the list can be resharded
        return super().state_dict()

    def load_state_dict(self, state_dicts, sharded_input=False):
        mine = super().load_state_dict(state_dicts, sharded_input)
        if self.g_state is not None:
            self.generator.set_state(self.g_state)
        self.buffer_size = len(self.buffer)
        return mine


class BufferDataset(_WrapperDataset):
    """Pack variable-length chunks into fixed ``seq_len`` lines.  ``pack_hard`` splits chunks across
    lines, otherwise lines are padded.  Optional per-line BOS/EOS (not duplicated when already
    present).  The residual buffer is plain state: dropped on rescale."""

    def __init__(self, dataset: _StatefulDataset, seq_len: int, pack_hard: bool, bos_token=None, eos_token=None,
                 pad_token=None):
        super().__init__(dataset)
        self.len = seq_len
        self.buffer: List[Any] = []
        self.bos, self.eos, self.pad = bos_token, eos_token, pad_token
        self.pack_hard = pack_hard
        if not pack_hard:
            assert pad_token is not None, "Error: if using pads, you must supply a pad_token"
        self.state_params = ["buffer"]

    def _cut(self, buffer: List[Any], length: int) -> Tuple[List[Any], List[Any]]:
        """Split off one output line; if an EOS must overwrite the last token, that token is carried over."""
        out, rest = buffer[:length], buffer[length:]
        if self.eos is not None and out[-1] != self.eos:
            rest = [out[-1]] + rest
            out[-1] = self.eos
        return out, rest

    def _get_buffer(self, iterable, length: int, buffer: List[Any]):
        new: List[Any] = []
        while len(buffer) + len(new) < length:  # pull until the next chunk would overrun the line
            buffer += new
            new = next(iterable)
        if self.bos is not None and (len(buffer) == 0 or buffer[0] != self.bos):
            buffer = [self.bos] + buffer
        if len(buffer) >= length:
            out, buffer = self._cut(buffer, length)
            buffer = buffer + new
        elif self.pack_hard:
            out, buffer = self._cut(buffer + new, length)
        else:
            if self.eos is not None and buffer[-1] != self.eos:
                buffer.append(self.eos)
            out = buffer + [self.pad] * (length - len(buffer)) if self.pad is not None else buffer
            buffer = new
        return out, buffer

    def __iter__(self):
        src = iter(self.dataset)
        while True:
            out, self.buffer = self._get_buffer(src, self.len, self.buffer)
            yield out


# -------------------------------------------------------------------------------------- base reader
class StreamingDocDataset(_StatefulDataset):
    """Distributed reader over a directory of shard files.

    Partitioning: every shard file is cut into ``worldsize`` fragments and rank r owns the contiguous
    fragment span ``[r*F, (r+1)*F)`` (F = number of files), i.e. a contiguous document range per file.
    Order: owned files are shuffled with ``random.seed(seed + rank)``; documents inside a file are
    visited through an LCG bijection (a=5, c=2(rank+seed)+1, m=2^ceil(log2 n), rejection sampling),
    so no index list is ever materialised.  Documents are emitted as chunks of at most
    ``max_chunksize`` tokens; the last chunk carries the delimiter, the first an optional BOS.
    Resumes mid-document and replays the skipped head chunks at the end of the epoch."""

    def __init__(self, datapath: str, rank: int, worldsize: int, filehandler: _ShardFileHandler, delimiter_token: Any,
                 bos_token: Optional[Any] = None, strip_tokens: Optional[Set[Any]] = None, seed: int = 42,
                 min_length: int = 1, max_chunksize: int = 1024, verbose: bool = False):
        super().__init__(datapath, rank, worldsize)
        assert max_chunksize > 0, "Max chunksize must be a nonzero positive integer"
        self.seed = seed
        self.filehandler = filehandler
        self.min_length = min_length
        self.chunksize = max_chunksize
        self.eos = delimiter_token
        self.bos = bos_token
        self.drop = set() if strip_tokens is None else strip_tokens
        self.verbose = verbose
        self.docset: List[Tuple[str, int, int]] = []  # (relative shard path, first doc, last doc) inclusive
        # position
        self.docset_index = 0
        self.chunk_index = -1
        # statistics
        self.epochs_seen = -1
        self.tokens_seen = 0
        self.docs_seen = 0
        self.percent_seen = 0
        self.state_params = ["dataset", "docset_index", "chunk_index", "epochs_seen", "tokens_seen", "docs_seen",
                             "percent_seen", "lcg_state"]
        self._len = 0
        self.dataset = ""
        self.lcg_state = 0

    # -- setup helpers
    def _list_shards(self) -> List[str]:
        root_len = len(self.datapath) + 1
        found = [os.path.join(root, name)[root_len:]
                 for root, _, files in os.walk(self.datapath, topdown=False) for name in files
                 if self.filehandler.is_legal(os.path.join(root, name))]
        found.sort()  # identical order on every machine
        return found

    def _doc_counts(self, pardir: str, dataset: str, shards_needed: Set[str]) -> Dict[str, int]:
        meta = os.path.join(pardir, "meta")
        countfiles = [f for f in os.listdir(meta) if "counts" in f and "csv" in f] if os.path.exists(meta) else []
        counts: Dict[str, int] = {}
        if countfiles:
            with open(os.path.join(meta, countfiles[0]), "r") as fh:
                for row in csv.DictReader(fh):
                    full = row["dataset/filename"]
                    at = full.find("/" + dataset) + 1
                    if at > 0:
                        counts[full[at + len(dataset) + 1:]] = int(row["documents"])
        else:
            counts = {s: self.filehandler.length(os.path.join(self.datapath, s)) for s in shards_needed}
        return counts

    def setup(self):
        if self.is_setup:
            return
        super().setup()
        head, tail = self.datapath, ""
        while len(tail) == 0:  # tolerate trailing slashes
            head, tail = os.path.split(head)
        pardir, self.dataset = head, tail

        shards = self._list_shards()
        W = self.worldsize
        first = (self.rank * W * len(shards)) // W
        last = ((self.rank + 1) * W * len(shards)) // W
        frags = [(shards[i // W], i % W) for i in range(first, last)]
        counts = self._doc_counts(pardir, self.dataset, {s for s, _ in frags})

        spans: Dict[str, List[int]] = {}
        for shard, frag in frags:
            n = counts[shard]
            lo = (n * frag) // W
            hi = (n * frag + n) // W - 1  # inclusive
            if shard not in spans:
                spans[shard] = [lo, hi]
            spans[shard][0] = min(spans[shard][0], lo)
            spans[shard][1] = max(spans[shard][1], hi)
        self.docset = [(s, lo, hi) for s, (lo, hi) in spans.items()]
        self._len = sum(hi - lo + 1 for _, lo, hi in self.docset)
        if self.verbose:
            logging.info(f"    Worker {self.rank} ingested {len(frags)} shard fragments from {self.dataset}")

        seed = self.seed + self.rank
        random.seed(seed)
        random.shuffle(self.docset)  # file order: different on every worker
        self.lcg_state = seed        # document order inside files: same guarantee

    # -- indexing
    def _get_docid(self, i: int):
        assert i <= self._len, f"You have requested an illegal doc index {i}, docset length is {self._len}"
        seen = 0
        for shard, lo, hi in self.docset:
            span = hi - lo + 1
            seen += span
            if seen > i:
                return shard, span, lo

    def _get_reader(self, path, newpath, reader):
        if newpath != path:
            del reader
            if self.verbose:
                logging.info(f"Worker {self.rank} opening new file {newpath}")
            reader = self.filehandler.open(newpath)
            path = newpath
        return path, reader

    def _random_map_docid(self, size: int) -> int:
        """Next index of the LCG permutation of range(size) after ``self.lcg_state``."""
        m = 2 ** math.ceil(math.log2(size))
        a, c = 5, (self.rank + self.seed) * 2 + 1
        state = self.lcg_state
        while True:
            state = (a * state + c) % m
            if state < size:
                return state

    def _construct_chunk(self, j: int, doc, n_chunks: int) -> List[Any]:
        start, n_pull = j * self.chunksize, self.chunksize
        if self.bos is not None:
            if j == 0:
                n_pull -= 1
            else:
                start -= 1
        chunk = self.filehandler.slice(doc, start, n_pull)
        self.tokens_seen += len(chunk)
        if self.bos is not None and j == 0:
            chunk = [self.bos] + chunk
        if j == n_chunks - 1:
            chunk = chunk + [self.eos]
        return chunk

    def _doc_len(self, doc) -> int:
        return len(doc) + (1 if self.bos is None else 2)

    def __iter__(self):
        if not self.is_setup:
            self.setup()
        start_doc = self.docset_index
        start_lcg = self.lcg_state
        skip_chunks = self.chunk_index + 1  # resume AFTER the last emitted chunk
        ndocs = self._len
        path, reader = "", None
        while True:
            for i in range(ndocs):
                doc_index = (start_doc + i) % ndocs
                if doc_index == 0:
                    self.epochs_seen += 1
                self.docset_index = doc_index
                shard, span, first = self._get_docid(doc_index)
                path, reader = self._get_reader(path, os.path.join(self.datapath, shard), reader)
                mapped = self._random_map_docid(span)
                doc = self.filehandler.get(reader, mapped + first, self.drop)
                if len(doc) == 0:
                    continue
                doclen = self._doc_len(doc)
                if doclen >= self.min_length:
                    n_chunks = math.ceil(doclen / self.chunksize)
                    for j in range(n_chunks):
                        if i == 0 and j < skip_chunks:
                            continue
                        self.chunk_index = j
                        if j == n_chunks - 1:
                            self.docs_seen += 1
                            self.percent_seen = self.docs_seen * 100 / (self._len + 1e-9)
                        yield self._construct_chunk(j, doc, n_chunks)
                self.lcg_state = mapped
            # epoch tail: the head chunks of the first document that the resume skipped
            self.docset_index = start_doc
            self.lcg_state = start_lcg
            shard, span, first = self._get_docid(start_doc)
            docid = self._random_map_docid(span) + first
            path, reader = self._get_reader(path, os.path.join(self.datapath, shard), reader)
            doc = self.filehandler.get(reader, docid, self.drop)
            if len(doc) == 0:
                continue
            doclen = self._doc_len(doc)
            if doclen >= self.min_length:
                n_chunks = math.ceil(doclen / self.chunksize)
                for j in range(skip_chunks):
                    self.chunk_index = j
                    yield self._construct_chunk(j, doc, n_chunks)

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        assert self.load_worldsize == self.worldsize, (
            f"StreamingDocDataset does not support rescaling (ckp size: {self.load_worldsize}, "
            f"world size: {self.worldsize}). Please use a ScalableShardDataset.")
        expected = self.dataset
        out = super().load_state_dict(state_dicts, sharded_input)
        assert expected == self.dataset, f"Dataset mismatch: checkpoint contains {self.dataset}, expected {expected}"
        return out


# ---------------------------------------------------------------------------------- rescalable layer
class ScalableShardDataset(_WrapperDataset):
    """Rescalability: the data is cut into ``n_logical_shards`` *logical* readers (a fixed number,
    independent of the job size); each physical worker owns ``n_logical_shards / worldsize`` of them,
    and draws the next document from a logical shard sampled proportionally to its documents
    remaining.  Logical-reader states are the unit that moves between workers on rescale."""

    def __init__(self, dataset: StreamingDocDataset, delimiter_token: Any, n_logical_shards: int = 2048,
                 verbose: bool = False):
        super().__init__(dataset)
        assert n_logical_shards % self.worldsize == 0, \
            f"World size {self.worldsize} must divide n_logical_shards {n_logical_shards} evenly"
        assert n_logical_shards > 0, f"n_logical_shards {n_logical_shards} must be a positive integer"
        self.total_shards = n_logical_shards
        self.delimiter = delimiter_token
        self.verbose = verbose
        self.data: List[StreamingDocDataset] = []
        self.logicals_owned: List[int] = []
        self.n_logicals = 0
        self.n_docs_remaining: List[int] = []
        self.generator = None
        # order-preserving state (only meaningful when the worker count is unchanged)
        self.current_reader = None
        self.logical_shard_states = None
        self.g_state = None
        self.state_params = ["current_reader", "g_state"]
        self.reshard_params = ["n_docs_remaining", "logical_shard_states"]

    def setup(self):
        if self.is_setup:
            return
        _StatefulDataset.setup(self)
        total = self.total_shards
        self.logicals_owned = _shard_partition(list(range(total)), self.rank, self.worldsize)
        self.n_logicals = total // self.worldsize
        assert len(self.logicals_owned) == self.n_logicals, \
            "(world size * num workers) does not divide logical shards evenly"
        for k, logical_rank in enumerate(self.logicals_owned):
            reader = deepcopy(self.dataset)
            reader.worldsize = reader.load_worldsize = total
            reader.rank = logical_rank
            reader.local_worldsize = 1
            reader.datapath = self.datapath
            reader.verbose = self.rank == 0
            self.data.append(reader)
            if self.verbose:
                logging.info(f"Worker {self.rank} assembled logical shard {logical_rank}, {k + 1} of {self.n_logicals}")
        for d in self.data:
            d.setup()
        self.n_docs_remaining = [d._len for d in self.data]
        self.generator = torch.Generator().manual_seed(self.rank)

    def __iter__(self):
        self.setup()
        streams = [iter(d) for d in self.data]
        while True:
            if self.current_reader is not None:
                ind = self.current_reader  # resume the document in flight
            else:
                assert sum(self.n_docs_remaining) > 0, f"No documents detected in {self.datapath}"
                ind = torch.multinomial(torch.tensor(self.n_docs_remaining, dtype=torch.float), 1,
                                        generator=self.generator).item()
            self.current_reader = ind
            out = next(streams[ind])
            while out[-1] != self.delimiter:  # whole documents only
                yield out
                out = next(streams[ind])
            self.current_reader = None
            self.n_docs_remaining[ind] -= 1
            if sum(self.n_docs_remaining) == 0:  # epoch boundary
                self.n_docs_remaining = [d._len for d in self.data]
                self.generator.manual_seed(self.rank)
            yield out

    def state_dict(self):
        self.setup()
        self.g_state = self.generator.get_state()
        self.logical_shard_states = [d.state_dict() for d in self.data]
        return _StatefulDataset.state_dict(self)

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        mine = _StatefulDataset.load_state_dict(self, state_dicts, sharded_input)
        if self.g_state is not None:
            self.generator.set_state(self.g_state)
        for i in range(self.n_logicals):
            self.data[i].load_state_dict([self.logical_shard_states[i]], True)
        return mine


# ------------------------------------------------------------------------------------ corpus mixing
class SamplingDataset(_WrapperDataset):
    """Mix several corpora (sub-directories of ``datapath``) by token share: always continue with the
    corpus whose emitted-token fraction is furthest below its target weight; whole documents only."""

    def __init__(self, datapath: str, dataset: Union[ScalableShardDataset, StreamingDocDataset], delimiter_token: Any,
                 datasets=None, weights=None, verbose=False):
        super().__init__(dataset)
        self.datapath = datapath
        self.delimiter = delimiter_token
        self.verbose = verbose
        if datasets is None:
            datasets = [f for f in os.listdir(datapath)
                        if not os.path.isfile(os.path.join(datapath, f)) and "meta" not in f]
        self.datasets = datasets
        assert len(self.datasets) > 0, "You must specify at least one dataset"
        if weights is not None:
            assert len(weights) == len(self.datasets), \
                f"Number of oversample weights {len(weights)} must match number of datasets {len(self.datasets)}"
            for w in weights:
                assert w > 0, f"Sampling rate {w} must be positive"
        raw = [1] * len(self.datasets) if weights is None else weights
        self.weights = [w / sum(raw) for w in raw]
        self.tokens_seen = [0] * len(self.datasets)
        self.current_iterator = -1
        self.state_params = ["tokens_seen", "current_iterator"]

    def setup(self):
        if self.is_setup:
            return
        _StatefulDataset.setup(self)
        self.data = []
        for k, name in enumerate(self.datasets):
            sub = deepcopy(self.dataset)
            sub.datapath = os.path.join(self.datapath, name)
            sub.rank, sub.worldsize, sub.local_worldsize = self.rank, self.worldsize, self.local_worldsize
            self.data.append(sub)
            if self.verbose:
                logging.info(f"Worker {self.rank} assembled subdataset iterator for {name}, {k + 1} of {len(self.datasets)}")
        for d in self.data:
            d.setup()

    def __iter__(self):
        self.setup()
        streams = [iter(d) for d in self.data]
        while True:
            if self.current_iterator != -1:
                out = next(streams[self.current_iterator])
                self.tokens_seen[self.current_iterator] += len(out)
                if out[-1] == self.delimiter:
                    self.current_iterator = -1
                yield out
            else:
                total = sum(self.tokens_seen) + 1e-9
                deficit = [self.weights[i] - self.tokens_seen[i] / total for i in range(len(self.datasets))]
                self.current_iterator = max((d, i) for i, d in enumerate(deficit))[1]

    def state_dict(self):
        self.setup()
        out = {self.statename("sample_iterator_states"): [d.state_dict() for d in self.data]}
        out.update(_StatefulDataset.state_dict(self))
        return out

    def load_state_dict(self, state_dicts, sharded_input=False):
        self.setup()
        mine = _StatefulDataset.load_state_dict(self, state_dicts, sharded_input)
        key = self.statename("sample_iterator_states")
        for i, sub in enumerate(self.data):
            sub.load_worldsize = self.load_worldsize
            sub.load_state_dict([sd[key][i] for sd in mine], True)
        return mine
