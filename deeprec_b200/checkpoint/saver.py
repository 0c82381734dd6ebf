"""Full + incremental checkpointing of dense state and EmbeddingVariables.

Format parity with the reference (docs/docs_en/Embedding-Variable-Export-Format.md):
per EV ``<name>-keys | -values | -freqs | -versions | -keys_filtered | -freqs_filtered | -versions_filtered |
-partition_offset | -partition_filter_offset`` with keys bucketed by ``key % 1000`` (1000 logical partitions =>
any N -> M re-shard on restore keeps ``key % 1000 % M == p``), optimizer slots as sibling groups
``<name>/<Slot>-keys/-values``, eviction executed at save time (single_tier_storage.h:235-261), filtered
(un-admitted) keys saved unless ``TF_EV_SAVE_FILTERED_FEATURES=0``, ``TF_EV_RESET_VERSION`` honoured on restore.
Incremental checkpoints (kernels/incr_save_restore_ops.h:347-500, python/training/incremental_saver.py:307-554):
only rows touched since the previous save (dirty bit set by the apply kernels) are written as
``-sparse_incr_keys/-values/-versions/-freqs/-incr_partition_offset``; restore = last full + later deltas.

Container: csrc/host/io_runtime.cc tensor bundle (8 MiB streaming writer, CRC32 per tensor, atomic publish).
"""
from __future__ import annotations

import ctypes as C
import json
import threading
import os
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional

import torch
from torch import nn

from .. import _native
from .._native import ptr
from ..config import env_flag
from ..embedding_variable import EmbeddingVariable
from ..optim.optimizers import DeepRecOptimizer, collect_embedding_variables

_DT = {torch.float32: "f32", torch.float64: "f64", torch.int64: "i64", torch.int32: "i32", torch.uint8: "u8",
       torch.bfloat16: "bf16", torch.float16: "f16", torch.int16: "i16", torch.int8: "i8", torch.bool: "b1"}
_RDT = {v: k for k, v in _DT.items()}


class BundleWriter:
    def __init__(self, prefix: str):
        os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
        self.lib = _native.host()
        self.h = self.lib.dr_bundle_writer_open(prefix.encode())
        if not self.h:
            raise IOError(f"cannot open checkpoint bundle {prefix}")

    def add(self, name: str, t: torch.Tensor) -> None:
        t = t.detach().to("cpu").contiguous()
        shape = (C.c_int64 * max(1, t.dim()))(*t.shape)
        rc = self.lib.dr_bundle_writer_add(self.h, name.encode(), _DT[t.dtype].encode(), shape, t.dim(), ptr(t) if t.numel() else None,
                                           t.numel() * t.element_size())
        if rc != 0:
            raise IOError(f"bundle write failed for {name}: {rc}")

    def close(self) -> None:
        rc = self.lib.dr_bundle_writer_close(self.h)
        self.h = None
        if rc != 0:
            raise IOError(f"bundle close failed: {rc}")


class BundleReader:
    def __init__(self, prefix: str):
        self.lib = _native.host()
        self.h = self.lib.dr_bundle_reader_open(prefix.encode())
        if not self.h:
            raise FileNotFoundError(f"no checkpoint bundle at {prefix}")
        self.entries: Dict[str, tuple] = {}
        name = C.create_string_buffer(1024); dt = C.create_string_buffer(16)
        shape = (C.c_int64 * 8)(); nb = C.c_int64(0)
        for i in range(self.lib.dr_bundle_reader_count(self.h)):
            nd = self.lib.dr_bundle_reader_entry(self.h, i, name, 1024, dt, 16, shape, C.byref(nb))
            self.entries[name.value.decode()] = (dt.value.decode(), tuple(shape[d] for d in range(nd)), nb.value)

    def has(self, name: str) -> bool:
        return name in self.entries

    def read(self, name: str, verify: bool = True) -> torch.Tensor:
        dt, shape, nb = self.entries[name]
        t = torch.empty(shape, dtype=_RDT[dt])
        rc = self.lib.dr_bundle_reader_read(self.h, name.encode(), ptr(t) if t.numel() else None, nb, int(verify))
        if rc != 0:
            raise IOError(f"bundle read failed for {name}: {rc} (-3 = CRC mismatch)")
        return t

    def close(self) -> None:
        if self.h:
            self.lib.dr_bundle_reader_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _state_file(directory: str, name: str = "checkpoint") -> str:
    return os.path.join(directory, name)


def _read_state(directory: str, name: str = "checkpoint") -> dict:
    p = _state_file(directory, name)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {"model_checkpoint_path": None, "all_model_checkpoint_paths": []}


def _write_state(directory: str, st: dict, name: str = "checkpoint") -> None:
    # unique tmp name: several shards / ranks may publish into the same directory at the same time
    tmp = f"{_state_file(directory, name)}.tmp.{os.getpid()}.{threading.get_ident()}"
    with open(tmp, "w") as f:
        json.dump(st, f)
    os.replace(tmp, _state_file(directory, name))


class _DirLock:
    """Advisory lock serialising the read-modify-write of a directory's state file across processes (sharded savers, PS shards)."""

    def __init__(self, directory: str):
        self.path = os.path.join(directory, ".checkpoint.lock")

    def __enter__(self):
        import fcntl
        self.f = open(self.path, "a+")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()
        return False


def _register_checkpoint(directory: str, prefix: str, step: int, base: str, max_to_keep: int) -> None:
    """Retention is per STEP, not per file: a step written as several shards (``<prefix>.ps{i}-<step>``, ``-partXXXXX-of-YYYYY``)
    is kept or evicted as a whole, so the shards of the step being written are never pruned by its own later shards."""
    with _DirLock(directory):
        st = _read_state(directory)
        ents = st.get("entries")
        if ents is None:
            ents = [{"path": q, "step": _step_of(q)} for q in st.get("all_model_checkpoint_paths", [])]
        ents = [e for e in ents if e["path"] != prefix] + [{"path": prefix, "step": int(step)}]
        steps = sorted({e["step"] for e in ents})
        evict = set(steps[:-max_to_keep]) if max_to_keep and len(steps) > max_to_keep else set()
        for e in ents:
            if e["step"] in evict:
                for ext in (".data", ".index"):
                    if os.path.exists(e["path"] + ext):
                        os.remove(e["path"] + ext)
        ents = [e for e in ents if e["step"] not in evict]
        st["entries"] = ents
        st["all_model_checkpoint_paths"] = [e["path"] for e in ents]
        st["model_checkpoint_path"] = prefix
        st.setdefault("latest_by_base", {})[base] = prefix
        _write_state(directory, st)


def _step_of(prefix: str) -> int:
    import re
    m = re.search(r"-(\d+)(?:-part\d+-of-\d+)?$", prefix)
    return int(m.group(1)) if m else -1


def latest_checkpoint(directory: str, base: Optional[str] = None) -> Optional[str]:
    """``tf.train.latest_checkpoint``.  ``base`` (the ``save_path`` basename a saver writes under, e.g. ``model.ps3``) selects one
    shard's newest checkpoint when several savers share the directory."""
    st = _read_state(directory)
    if base is not None:
        return st.get("latest_by_base", {}).get(base)
    return st.get("model_checkpoint_path")


class Saver:
    """``tf.train.Saver(sharded=..., incremental_save_restore=...)`` for a module + its EmbeddingVariables + optimizer."""

    def __init__(self, module: Optional[nn.Module] = None, embedding_variables: Optional[Iterable[EmbeddingVariable]] = None,
                 optimizer: Optional[DeepRecOptimizer] = None, extra_state: Optional[Dict[str, torch.Tensor]] = None,
                 max_to_keep: int = 5, sharded: bool = False, partition_id: int = 0, partition_num: int = 1,
                 async_restore_threads: int = 4):
        self.module, self.optimizer, self.extra = module, optimizer, extra_state or {}
        evs = list(embedding_variables) if embedding_variables is not None else (collect_embedding_variables(module) if module is not None else [])
        if optimizer is not None:
            seen = {id(e) for e in evs}
            evs += [e for e in optimizer.evs if id(e) not in seen]
        self.evs: List[EmbeddingVariable] = evs
        self.max_to_keep, self.sharded = max_to_keep, sharded
        self.partition_id, self.partition_num = partition_id, partition_num
        self.restore_threads = async_restore_threads
        self.save_filtered = env_flag("TF_EV_SAVE_FILTERED_FEATURES", True)

    # ------------------------------------------------------------------------------------------------
    def _dense_tensors(self) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        anchors = {id(e._anchor) for e in self.evs}
        if self.module is not None:
            for k, v in self.module.state_dict().items():
                if v.numel() == 0 or id(v) in anchors or k.endswith("_anchor"):
                    continue
                out["dense/" + k] = v
        if self.optimizer is not None:
            names = {}
            if self.module is not None:
                names = {id(p): n for n, p in self.module.named_parameters()}
            for p, st in self.optimizer.state.items():
                base = names.get(id(p), f"param{len(out)}")
                for sk, sv in st.items():
                    out[f"opt/{base}/{sk}"] = sv if torch.is_tensor(sv) else torch.tensor(sv)
            out["opt/beta1_power"] = torch.tensor(float(self.optimizer.beta1_power), dtype=torch.float64)
            out["opt/beta2_power"] = torch.tensor(float(self.optimizer.beta2_power), dtype=torch.float64)
        for k, v in self.extra.items():
            out["extra/" + k] = v
        return out

    def _write_ev(self, w: BundleWriter, ev: EmbeddingVariable, snap: Dict[str, torch.Tensor], incremental: bool) -> None:
        name, dim = ev.name, ev.embedding_dim
        rows = snap["rows"]
        if incremental:
            w.add(f"{name}-sparse_incr_keys", snap["keys"]); w.add(f"{name}-sparse_incr_values", rows[:, :dim].contiguous())
            w.add(f"{name}-sparse_incr_versions", snap["versions"]); w.add(f"{name}-sparse_incr_freqs", snap["freqs"])
            w.add(f"{name}-incr_partition_offset", snap["partition_offset"])
            for i, sn in enumerate(ev._slot_names):
                w.add(f"{name}/{sn}-sparse_incr_values", rows[:, (1 + i) * dim:(2 + i) * dim].contiguous())
        else:
            w.add(f"{name}-keys", snap["keys"]); w.add(f"{name}-values", rows[:, :dim].contiguous())
            w.add(f"{name}-freqs", snap["freqs"]); w.add(f"{name}-versions", snap["versions"])
            w.add(f"{name}-partition_offset", snap["partition_offset"])
            for i, sn in enumerate(ev._slot_names):
                w.add(f"{name}/{sn}-keys", snap["keys"]); w.add(f"{name}/{sn}-values", rows[:, (1 + i) * dim:(2 + i) * dim].contiguous())
        if ev._has_scalars:
            w.add(f"{name}-row_scalars" + ("_incr" if incremental else ""), rows[:, (1 + ev._num_slots) * dim:].contiguous())
        if self.save_filtered:
            sfx = "_incr" if incremental else ""
            w.add(f"{name}-keys_filtered{sfx}", snap["keys_filtered"]); w.add(f"{name}-freqs_filtered{sfx}", snap["freqs_filtered"])
            w.add(f"{name}-versions_filtered{sfx}", snap["versions_filtered"])
            w.add(f"{name}-partition_filter_offset{sfx}", snap["partition_filter_offset"])
        bloom = ev.table.bloom_state()
        if bloom is not None:
            w.add(f"{name}-bloom", bloom)

    def save(self, save_path: str, global_step: Optional[int] = None, incremental: bool = False) -> str:
        step = int(global_step if global_step is not None else (self.optimizer.global_step if self.optimizer else 0))
        prefix = f"{save_path}-{step}"
        if self.sharded and self.partition_num > 1:
            prefix += f"-part{self.partition_id:05d}-of-{self.partition_num:05d}"
        w = BundleWriter(prefix)
        w.add("global_step", torch.tensor(step, dtype=torch.int64))
        w.add("meta/ev_names", torch.tensor(list("\n".join(e.name for e in self.evs).encode()), dtype=torch.uint8))
        for k, v in self._dense_tensors().items():
            w.add(k, v)
        for ev in self.evs:
            if not incremental:
                ev.table.shrink(step)                     # eviction runs only inside a full save
            snap = ev.table.snapshot(dirty_only=incremental)
            self._write_ev(w, ev, snap, incremental)
        w.close()
        for ev in self.evs:
            ev.table.clear_dirty()                        # recorder is cleared on each (full or incremental) save
        d = os.path.dirname(os.path.abspath(save_path))
        if not incremental:
            _register_checkpoint(d, prefix, step, os.path.basename(save_path), self.max_to_keep)
        return prefix

    # ------------------------------------------------------------------------------------------------
    def _restore_ev(self, r: BundleReader, ev: EmbeddingVariable, incremental: bool, reset_version: bool) -> int:
        name, dim = ev.name, ev.embedding_dim
        kk, vv = (f"{name}-sparse_incr_keys", f"{name}-sparse_incr_values") if incremental else (f"{name}-keys", f"{name}-values")
        if not r.has(kk):
            return 0
        keys = r.read(kk)
        t = ev.table
        stride = t.stride
        rows = torch.zeros(keys.numel(), stride, dtype=torch.float32)
        rows[:, :dim] = r.read(vv)
        for i, sn in enumerate(ev._slot_names):
            sk = f"{name}/{sn}-sparse_incr_values" if incremental else f"{name}/{sn}-values"
            rows[:, (1 + i) * dim:(2 + i) * dim] = r.read(sk) if r.has(sk) else float(ev._slot_init[i])
        sc = f"{name}-row_scalars" + ("_incr" if incremental else "")
        if ev._has_scalars and r.has(sc):
            rows[:, (1 + ev._num_slots) * dim:] = r.read(sc)
        freqs = r.read(f"{name}-sparse_incr_freqs" if incremental else f"{name}-freqs")
        vers = r.read(f"{name}-sparse_incr_versions" if incremental else f"{name}-versions")
        n = t.import_(keys, rows, freqs, vers, self.partition_id, self.partition_num, reset_version)
        sfx = "_incr" if incremental else ""
        if r.has(f"{name}-keys_filtered{sfx}"):
            fk = r.read(f"{name}-keys_filtered{sfx}")
            if fk.numel():
                t.import_(fk, None, r.read(f"{name}-freqs_filtered{sfx}"), r.read(f"{name}-versions_filtered{sfx}"),
                          self.partition_id, self.partition_num, reset_version)
        if r.has(f"{name}-bloom"):
            t.load_bloom_state(r.read(f"{name}-bloom"))
        return n

    def restore(self, prefix: str, incremental: bool = False, strict: bool = True) -> int:
        """Restore dense state + every EV (EVs in parallel on a restore pool -- KvResourceImportV3 is async in the
        reference, kv_variable_restore_ops.cc:338).  Returns the checkpoint's global step."""
        r = BundleReader(prefix)
        step = int(r.read("global_step"))
        reset_version = env_flag("TF_EV_RESET_VERSION", False)
        if self.module is not None:
            sd = self.module.state_dict()
            for k in list(sd.keys()):
                if r.has("dense/" + k):
                    sd[k].copy_(r.read("dense/" + k).to(sd[k].device))
                elif strict and sd[k].numel() and not k.endswith("_anchor"):
                    raise KeyError(f"tensor dense/{k} missing from {prefix}")
        if self.optimizer is not None:
            names = {id(p): n for n, p in self.module.named_parameters()} if self.module is not None else {}
            for g in self.optimizer.param_groups:
                for p in g["params"]:
                    base = names.get(id(p))
                    if base is None:
                        continue
                    for key in [e for e in r.entries if e.startswith(f"opt/{base}/")]:
                        sk = key.rsplit("/", 1)[1]
                        v = r.read(key)
                        self.optimizer.state[p][sk] = v.to(p.device) if v.dim() else v.item()
            if r.has("opt/beta1_power"):
                self.optimizer.beta1_power = float(r.read("opt/beta1_power")); self.optimizer.beta2_power = float(r.read("opt/beta2_power"))
            self.optimizer.global_step.value = step
        for k in self.extra:
            if r.has("extra/" + k):
                self.extra[k].copy_(r.read("extra/" + k).to(self.extra[k].device))
        host_evs = [e for e in self.evs if e.device.type == "cpu"]
        dev_evs = [e for e in self.evs if e.device.type != "cpu"]
        if host_evs and self.restore_threads > 1:
            with ThreadPoolExecutor(self.restore_threads) as ex:
                list(ex.map(lambda e: self._restore_ev(BundleReader(prefix), e, incremental, reset_version), host_evs))
        else:
            for e in host_evs:
                self._restore_ev(r, e, incremental, reset_version)
        for e in dev_evs:                                   # device imports are stream-ordered kernels
            self._restore_ev(r, e, incremental, reset_version)
        for e in self.evs:
            e.table.clear_dirty()
        r.close()
        return step


class IncrementalSaver(Saver):
    """``tf.train.Saver(incremental_save_restore=True)`` + IncrementalSaver (incremental_saver.py:307-554): keeps a
    chain of delta checkpoints under ``<dir>/.incr`` and replays them over the last full checkpoint on restore."""

    def incremental_save(self, save_path: str, global_step: Optional[int] = None) -> str:
        d = os.path.dirname(os.path.abspath(save_path))
        incr_dir = os.path.join(d, ".incr")
        os.makedirs(incr_dir, exist_ok=True)
        prefix = super().save(os.path.join(incr_dir, os.path.basename(save_path)), global_step, incremental=True)
        st = _read_state(incr_dir, "incr_checkpoint")
        st.setdefault("chain", []).append({"path": prefix, "base": latest_checkpoint(d), "time": time.time()})
        st["model_checkpoint_path"] = prefix
        _write_state(incr_dir, st, "incr_checkpoint")
        return prefix

    def save(self, save_path: str, global_step: Optional[int] = None, incremental: bool = False) -> str:
        prefix = super().save(save_path, global_step, incremental)
        if not incremental:
            # a new full checkpoint starts a new delta chain
            d = os.path.dirname(os.path.abspath(save_path))
            incr_dir = os.path.join(d, ".incr")
            if os.path.isdir(incr_dir):
                st = _read_state(incr_dir, "incr_checkpoint")
                st["chain"] = [c for c in st.get("chain", []) if c.get("base") == prefix]
                _write_state(incr_dir, st, "incr_checkpoint")
        return prefix

    def recover_incr_checkpoints(self, directory: str) -> int:
        """Restore the latest full checkpoint, then replay newer incremental checkpoints in order."""
        full = latest_checkpoint(directory)
        if full is None:
            raise FileNotFoundError(f"no full checkpoint under {directory}")
        step = self.restore(full)
        incr_dir = os.path.join(directory, ".incr")
        st = _read_state(incr_dir, "incr_checkpoint") if os.path.isdir(incr_dir) else {}
        for c in st.get("chain", []):
            if c.get("base") == full and os.path.exists(c["path"] + ".index"):
                step = self.restore(c["path"], incremental=True, strict=False)
        return step
