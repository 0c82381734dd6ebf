"""Export a trained model for the native serving runtime (``CollectiveStrategy.export_saved_model`` / SavedModel analogue).

Layout:  <dir>/saved_model.json                      architecture + version + bundle path
         <dir>/variables/variables.{data,index}      tensor bundle: dense params, BatchNorm moving statistics, per-table
                                                     keys / values / freqs / versions / default matrix
         <root>/serving_versions.json                version file polled by the ModelUpdater (full + delta versions)
Delta exports (``export_delta``) contain only rows touched since the previous export (dirty bits) + the dense block."""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

from ..checkpoint.saver import BundleWriter


def _write_versions(root: str, full: Optional[dict] = None, delta: Optional[dict] = None) -> None:
    p = os.path.join(root, "serving_versions.json")
    st = {"full": None, "deltas": []}
    if os.path.exists(p):
        with open(p) as f:
            st = json.load(f)
    if full is not None:
        st["full"], st["deltas"] = full, []
    if delta is not None:
        st["deltas"].append(delta)
    tmp = p + ".tmp"
    with open(tmp, "w") as f:
        json.dump(st, f)
    os.replace(tmp, p)


def export_saved_model(engine, export_dir: str, version: Optional[int] = None, root: Optional[str] = None) -> str:
    """Full export of a :class:`DLRMEngine` (rank-local tables; rank 0 also writes the dense block)."""
    cfg = engine.cfg
    version = int(version if version is not None else engine.global_step())
    os.makedirs(os.path.join(export_dir, "variables"), exist_ok=True)
    w = BundleWriter(os.path.join(export_dir, "variables", "variables"))
    for name, (o, n) in engine.views.items():
        w.add("dense/" + name, engine.params[o:o + n])
    for L in engine.bot:
        w.add(f"bn/{L.name}/moving_mean", L.running_mean)
        w.add(f"bn/{L.name}/moving_variance", L.running_var)
    for t, tbl in engine.tables.items():
        s = tbl.snapshot()
        w.add(f"table/{t}-keys", s["keys"]); w.add(f"table/{t}-values", s["rows"][:, : engine.D].contiguous())
        w.add(f"table/{t}-freqs", s["freqs"]); w.add(f"table/{t}-versions", s["versions"])
        w.add(f"table/{t}-default", tbl.default_matrix)
        tbl.clear_dirty()
    w.close()
    meta = {"model": "dlrm", "version": version, "num_dense": cfg.num_dense, "num_tables": engine.T, "embedding_dim": engine.D,
            "mlp_bot": list(cfg.mlp_bot), "mlp_top": list(cfg.mlp_top), "bn_eps": cfg.bn_eps, "variables": "variables/variables",
            "signature": {"inputs": {"dense": ["B", cfg.num_dense], "ids": [engine.T, "B"]}, "outputs": {"probabilities": ["B"]}}}
    with open(os.path.join(export_dir, "saved_model.json"), "w") as f:
        json.dump(meta, f)
    _write_versions(root or export_dir, full={"version": version, "dir": os.path.abspath(export_dir)})
    return export_dir


def export_delta(engine, root: str, base_version: int, version: Optional[int] = None) -> str:
    """Incremental export: rows touched since the last export + the (small) dense block."""
    version = int(version if version is not None else engine.global_step())
    d = os.path.join(root, ".incr")
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, f"delta-{version}")
    w = BundleWriter(prefix)
    for name, (o, n) in engine.views.items():
        w.add("dense/" + name, engine.params[o:o + n])
    for L in engine.bot:
        w.add(f"bn/{L.name}/moving_mean", L.running_mean)
        w.add(f"bn/{L.name}/moving_variance", L.running_var)
    for t, tbl in engine.tables.items():
        s = tbl.snapshot(dirty_only=True)
        w.add(f"table/{t}-sparse_incr_keys", s["keys"]); w.add(f"table/{t}-sparse_incr_values", s["rows"][:, : engine.D].contiguous())
        tbl.clear_dirty()
    w.close()
    _write_versions(root, delta={"version": version, "base": int(base_version), "prefix": os.path.abspath(prefix)})
    return prefix
