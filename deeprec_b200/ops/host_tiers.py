"""Host-side multi-tier storage: DRAM (HostEV) over a log-structured SSD store -- ``StorageType.DRAM_SSDHASH`` and the lower two
tiers of ``HBM_DRAM_SSDHASH``.

Reference: framework/embedding/dram_ssd_storage.h, hbm_dram_ssd_storage.h, ssd_hash_kv.h:139-810.  DRAM holds at
most ``storage_size[dram tier] / row_bytes`` rows; after every mutating call the coldest rows (LFU: lowest frequency, LRU:
oldest version) are demoted to the SSD store; a lookup / apply that touches a key living only on SSD promotes it first, so all
compute runs against DRAM rows.  Checkpoints merge both tiers (rows on SSD are exported through the index, not by copying
.emb files), eviction policies at save time cover both tiers.
"""
from __future__ import annotations

import os
import tempfile
from typing import Dict, Optional

import torch

from .. import _native
from .._native import EvConfig, OptHyper, ptr
from ..embedding_variable import HostTable, _i64


class SsdStore:
    """ctypes handle on dr::SsdHashStore."""

    def __init__(self, stride: int, path: Optional[str] = None, file_bytes: int = 64 << 20, async_compaction: Optional[bool] = None):
        self.lib = _native.host()
        self.stride = int(stride)
        if async_compaction is None:      # TF_SSDHASH_ASYNC_COMPACTION in the reference
            async_compaction = os.environ.get("DEEPREC_SSDHASH_ASYNC_COMPACTION", os.environ.get("TF_SSDHASH_ASYNC_COMPACTION", "0")) == "1"
        base = path or tempfile.gettempdir()
        os.makedirs(base, exist_ok=True)
        self.dir = tempfile.mkdtemp(prefix="deeprec_ssd_", dir=base)
        self.h = self.lib.dr_ssd_create(self.dir.encode(), self.stride, int(file_bytes), int(bool(async_compaction)))

    def __del__(self):
        try:
            if self.h:
                self.lib.dr_ssd_destroy(self.h); self.h = None
        except Exception:
            pass

    def size(self) -> int:
        return int(self.lib.dr_ssd_size(self.h))

    def put(self, keys, rows, freqs, versions) -> None:
        k = _i64(keys).view(-1)
        if k.numel() == 0:
            return
        r = rows.to(torch.float32).contiguous()
        assert r.shape == (k.numel(), self.stride)
        self.lib.dr_ssd_put(self.h, ptr(k), ptr(r), ptr(_i64(freqs)), ptr(_i64(versions)), k.numel())

    def get(self, keys):
        k = _i64(keys).view(-1)
        n = k.numel()
        rows = torch.zeros(n, self.stride); f = torch.zeros(n, dtype=torch.int64); v = torch.full((n,), -1, dtype=torch.int64)
        found = torch.zeros(n, dtype=torch.uint8)
        if n:
            self.lib.dr_ssd_get(self.h, ptr(k), n, ptr(rows), ptr(f), ptr(v), ptr(found))
        return rows, f, v, found.bool()

    def contains(self, keys) -> torch.Tensor:
        k = _i64(keys).view(-1)
        found = torch.zeros(k.numel(), dtype=torch.uint8)
        if k.numel():
            self.lib.dr_ssd_contains(self.h, ptr(k), k.numel(), ptr(found))
        return found.bool()

    def remove(self, keys) -> int:
        k = _i64(keys).view(-1)
        return int(self.lib.dr_ssd_remove(self.h, ptr(k), k.numel())) if k.numel() else 0

    def keys(self) -> torch.Tensor:
        n = self.size()
        out = torch.empty(n + 1024, dtype=torch.int64)
        m = int(self.lib.dr_ssd_export_keys(self.h, ptr(out), out.numel()))
        return out[: min(m, out.numel())].clone()

    def compact(self, ratio: float = 0.5) -> int:
        return int(self.lib.dr_ssd_compact(self.h, float(ratio)))

    def stats(self) -> Dict[str, int]:
        return {"keys": self.size(), "files": int(self.lib.dr_ssd_num_files(self.h)), "bytes": int(self.lib.dr_ssd_bytes(self.h)),
                "compactions": int(self.lib.dr_ssd_compactions(self.h))}


class DramSsdTable:
    """HostTable interface over two tiers.  ``dram_rows`` = capacity of the DRAM tier in rows."""

    def __init__(self, cfg: EvConfig, default_matrix: torch.Tensor, dram_rows: int, path: Optional[str] = None, strategy: int = 0,
                 file_bytes: int = 64 << 20):
        self.cfg = cfg
        self.dram = HostTable(cfg, default_matrix)
        self.lib, self.h = self.dram.lib, self.dram.h          # callers that reach for the raw host handle see the DRAM tier
        self.dim, self.stride = self.dram.dim, self.dram.stride
        self.ssd = SsdStore(self.stride, path, file_bytes)
        self.dram_rows = max(64, int(dram_rows))
        self.strategy = int(strategy)
        self.device = torch.device("cpu")
        self.promotions = 0
        self.demotions = 0
        self._dirty_demoted: list = []          # keys demoted since the last clear_dirty(): still part of the next delta

    # ---- tier movement --------------------------------------------------------------------------------------------------
    def _promote(self, keys: torch.Tensor) -> None:
        """Move the listed keys SSD -> DRAM when they are not DRAM-resident."""
        if self.ssd.size() == 0:
            return
        k = torch.unique(_i64(keys).view(-1))
        on_ssd = self.ssd.contains(k)
        if not on_ssd.any():
            return
        k = k[on_ssd]
        rows, f, v, found = self.ssd.get(k)
        k, rows, f, v = k[found], rows[found], f[found], v[found]
        if k.numel():
            self.dram.import_(k, rows, f, v)
            self.ssd.remove(k)
            self.promotions += int(k.numel())

    def _demote(self, protect: Optional[torch.Tensor] = None) -> None:
        over = self.dram.size() - self.dram_rows
        if over <= 0:
            return
        snap = self.dram.snapshot()
        keys = snap["keys"]
        score = (snap["freqs"] if self.strategy == 0 else snap["versions"]).to(torch.float64)
        if protect is not None and protect.numel():
            score = torch.where(torch.isin(keys, _i64(protect).view(-1)), torch.full_like(score, float("inf")), score)
        n = min(int(keys.numel()), over + max(64, self.dram_rows // 16))
        n = min(n, int(torch.isfinite(score).sum()))
        if n <= 0:
            return
        idx = torch.topk(score, n, largest=False).indices
        self.ssd.put(keys[idx], snap["rows"][idx], snap["freqs"][idx], snap["versions"][idx])
        self._dirty_demoted.append(keys[idx].clone())
        self.dram.remove(keys[idx])
        self.demotions += n

    # ---- table interface ----------------------------------------------------------------------------------------------------
    def size(self) -> int:
        return self.dram.size() + self.ssd.size()

    def total_keys(self) -> int:
        return self.dram.total_keys() + self.ssd.size()

    def lookup(self, keys: torch.Tensor) -> torch.Tensor:
        self._promote(keys)
        out = self.dram.lookup(keys)
        self._demote(protect=keys)
        return out

    def lookup_slot(self, keys, slot):
        self._promote(keys)
        return self.dram.lookup_slot(keys, slot)

    def get_freq(self, keys):
        f = self.dram.get_freq(keys)
        if self.ssd.size():
            _, sf, _, found = self.ssd.get(keys)
            f = torch.where(found, sf, f)
        return f

    def get_version(self, keys):
        v = self.dram.get_version(keys)
        if self.ssd.size():
            _, _, sv, found = self.ssd.get(keys)
            v = torch.where(found, sv, v)
        return v

    def apply(self, keys, grads, counts, hp: OptHyper) -> None:
        self._promote(keys)
        self.dram.apply(keys, grads, counts, hp)
        self._demote(protect=keys)

    def apply_raw(self, ids, grads, hp: OptHyper) -> None:
        self._promote(ids)
        self.dram.apply_raw(ids, grads, hp)
        self._demote(protect=ids)

    def lookup_tier(self, keys: torch.Tensor) -> torch.Tensor:
        """0 = DRAM, 1 = SSD, -1 = absent."""
        k = _i64(keys).view(-1)
        n = k.numel()
        rows = torch.empty(n, self.stride); f = torch.empty(n, dtype=torch.int64); v = torch.empty(n, dtype=torch.int64)
        found = torch.zeros(n, dtype=torch.uint8)
        self.lib.dr_host_ev_export_keys(self.h, ptr(k), n, ptr(rows), ptr(f), ptr(v), ptr(found))
        out = torch.full((n,), -1, dtype=torch.int64)
        out[self.ssd.contains(k)] = 1
        out[found.bool()] = 0
        return out.view(keys.shape)

    def shrink(self, step: int) -> int:
        n = self.dram.shrink(step)
        # SSD tier: same predicates, evaluated on the exported rows
        sk = self.ssd.keys()
        if sk.numel() and (self.cfg.steps_to_live > 0 or self.cfg.l2_weight_threshold >= 0):
            rows, _, v, found = self.ssd.get(sk)
            dead = torch.zeros(sk.numel(), dtype=torch.bool)
            if self.cfg.steps_to_live > 0:
                dead |= (v >= 0) & (step - v > self.cfg.steps_to_live)
            if self.cfg.l2_weight_threshold >= 0:
                dead |= 0.5 * (rows[:, : self.dim] ** 2).sum(1) < self.cfg.l2_weight_threshold
            dead &= found
            n += self.ssd.remove(sk[dead])
        self.ssd.compact(0.5)
        return n

    def remove(self, keys) -> int:
        return self.dram.remove(keys) + self.ssd.remove(keys)

    def clear_dirty(self) -> None:
        self.dram.clear_dirty()
        self._dirty_demoted = []

    def export_keys(self, keys: torch.Tensor):
        """(rows [n, stride], freqs, versions, found) of specific keys -- SSD-resident keys are promoted first."""
        self._promote(keys)
        return self.dram.export_keys(keys)

    def snapshot(self, dirty_only: bool = False, part_id: int = 0, part_num: int = 1) -> Dict[str, torch.Tensor]:
        d = self.dram.snapshot(dirty_only, part_id, part_num)
        if self.ssd.size() == 0:
            return d
        if dirty_only:                      # only rows demoted since the last delta can be dirty on SSD
            if not self._dirty_demoted:
                return d
            sk = torch.unique(torch.cat(self._dirty_demoted))
        else:
            sk = self.ssd.keys()
        if part_num > 1:
            sk = sk[torch.remainder(torch.remainder(sk, 1000), part_num) == part_id]
        rows, f, v, found = self.ssd.get(sk)
        keys = torch.cat([d["keys"], sk[found]])
        rows = torch.cat([d["rows"], rows[found]]); f = torch.cat([d["freqs"], f[found]]); v = torch.cat([d["versions"], v[found]])
        b = torch.remainder(keys, 1000)
        o = torch.argsort(keys, stable=True); o = o[torch.argsort(b[o], stable=True)]
        off = torch.zeros(1001, dtype=torch.int64); off[1:] = torch.cumsum(torch.bincount(b, minlength=1000), 0)
        out = dict(d)
        out.update(keys=keys[o], rows=rows[o], freqs=f[o], versions=v[o], partition_offset=off)
        return out

    def import_(self, keys, rows, freqs, versions, part_id=0, part_num=1, reset_version=False) -> int:
        if self.ssd.size():                 # imported rows supersede any (stale) copy on the SSD tier
            k = _i64(keys).view(-1)
            if part_num > 1:
                k = k[torch.remainder(torch.remainder(k, 1000), part_num) == part_id]
            self.ssd.remove(k)
        n = self.dram.import_(keys, rows, freqs, versions, part_id, part_num, reset_version)
        self._demote()
        return n

    def bloom_state(self):
        return self.dram.bloom_state()

    def load_bloom_state(self, s):
        self.dram.load_bloom_state(s)

    def tier_stats(self) -> Dict[str, int]:
        return {"dram_rows": self.dram.size(), "ssd": self.ssd.stats(), "promotions": self.promotions, "demotions": self.demotions}
