"""Post-training low-precision optimisation of a served model -- ``optimize(model_path, save_path, opt_config=None, data_type="BF16", calib_data=None)``
and ``convert_ckpt(ckpt_prefix, save_prefix, opt_model_path)``, the interface of the reference's ``tools/low_precision_optimize`` (BF16 / FP16 / INT8 of
a SavedModel, embeddings included, per-node overrides, a calibration set, and the same decisions re-applied to a later checkpoint).

What it does here: a saved-model directory written by ``serving.export`` (DLRM module export or any op-program export) is rewritten with

* every embedding table's rows (``table/<t>-values``, ``-sparse_incr_values`` of deltas, and its default-value matrix) as INT8 with one fp32 scale
  per row, or BF16 / FP16;
* every dense kernel (``prog/<op>/kernel``, ``dense/.../kernel``, GRU / attention matrices) as BF16 / FP16, or INT8 with one scale per output row;
  biases and BatchNorm / LayerNorm vectors stay fp32;

and a ``"low_precision"`` record in ``saved_model.json``.  Both native Processors read float tensors through ``dr::ReadAsFloat`` (csrc/common/bundle.h),
so an optimised model (2-4x smaller to ship; the same for every delta after ``convert_ckpt``) loads and updates like the original.  ``calib_data``
(a list of ``(dense, ids)`` request arrays) is used the way an offline check should be: the original and the optimised model both serve it on the
CPU Processor and the largest probability drift is reported (and bounded with ``max_drift``).

``convert(prefix_in, prefix_out, dtype)`` keeps the plain checkpoint-bundle form (training checkpoints: dense tensors and EV ``-values``)."""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
from typing import Dict, Optional, Sequence

import torch

from ..checkpoint.saver import BundleReader, BundleWriter

_TYPES = {"BF16": "bf16", "FP16": "fp16", "INT8": "int8", "FP32": "fp32"}


def _quant_int8(t: torch.Tensor):
    t = t.float()
    if t.dim() >= 2:
        scale = t.abs().amax(dim=tuple(range(1, t.dim()))).clamp_min(1e-12) / 127.0           # one scale per leading-dimension row
        q = torch.clamp(torch.round(t / scale.view(-1, *([1] * (t.dim() - 1)))), -127, 127).to(torch.int8)
    else:
        scale = (t.abs().max().clamp_min(1e-12) / 127.0).reshape(1)
        q = torch.clamp(torch.round(t / scale), -127, 127).to(torch.int8)
    return q, scale.float().contiguous()


def _store(w: BundleWriter, name: str, t: torch.Tensor, dtype: str) -> int:
    if dtype in ("bf16", "fp16"):
        q = t.to(torch.bfloat16 if dtype == "bf16" else torch.float16)
        w.add(name, q)
        return q.numel() * 2
    if dtype == "int8":
        q, s = _quant_int8(t)
        w.add(name, q); w.add(name + "/scale", s)
        return q.numel() + s.numel() * 4
    w.add(name, t)
    return t.numel() * t.element_size()


def _node_of(name: str) -> Optional[tuple]:
    """(node name as users address it, kind) of an optimisable tensor, or None."""
    if name.startswith("table/") and (name.endswith("-values") or name.endswith("-sparse_incr_values") or name.endswith("-default")):
        return name.split("-")[0], "embedding"                                                # table/<t>
    if name.endswith("/kernel") or name.rsplit("/", 1)[-1] in ("w_ih", "w_hh", "w1", "w2"):
        return name.rsplit("/", 1)[0] if name.endswith("/kernel") else name, "dense"
    return None


def _plan(entries, opt_config: Optional[Dict[str, str]], data_type: str) -> Dict[str, str]:
    """tensor name -> storage dtype.  With ``opt_config`` only the named nodes are optimised (the reference's contract), else every optimisable node."""
    default = _TYPES[data_type.upper()]
    cfg = None if opt_config is None else {k: _TYPES[v.upper()] for k, v in opt_config.items()}
    plan = {}
    for name, (dt, shape, _) in entries.items():
        node = _node_of(name)
        if node is None or dt != "f32":
            continue
        dtype = default if cfg is None else cfg.get(node[0], cfg.get(name))
        if dtype and dtype != "fp32":
            plan[name] = dtype
    if cfg is not None:
        known = {(_node_of(n) or (None,))[0] for n in entries} | set(entries)
        unknown = [k for k in cfg if k not in known]
        if unknown:
            raise KeyError(f"opt_config names nodes the model does not have: {unknown}")
    return plan


def _rewrite(prefix_in: str, prefix_out: str, plan: Dict[str, str], log=print) -> dict:
    r = BundleReader(prefix_in)
    w = BundleWriter(prefix_out)
    before = after = 0
    for name, (dt, shape, nbytes) in r.entries.items():
        t = r.read(name)
        before += nbytes
        if name in plan:
            after += _store(w, name, t, plan[name])
            if log:
                log(f"Optimize {_node_of(name)[1]} to {plan[name].upper()}: {name}")
        else:
            w.add(name, t); after += nbytes
    w.close(); r.close()
    return {"bytes_before": before, "bytes_after": after, "ratio": after / max(1, before)}


def optimize(model_path: str, save_path: str, opt_config: Optional[Dict[str, str]] = None, data_type: str = "BF16",
             calib_data: Optional[Sequence] = None, max_drift: Optional[float] = None, log=print) -> dict:
    """Rewrite the saved model under ``model_path`` into ``save_path`` in low precision; returns sizes, the per-tensor plan and (with ``calib_data``)
    the largest probability drift against the original on the CPU Processor."""
    with open(os.path.join(model_path, "saved_model.json")) as f:
        meta = json.load(f)
    var = meta.get("variables", "variables/variables")
    src = os.path.join(model_path, var)
    r = BundleReader(src)
    plan = _plan(r.entries, opt_config, data_type)
    r.close()
    os.makedirs(os.path.join(save_path, os.path.dirname(var)), exist_ok=True)
    if log:
        log("Optimization Result:")
    res = _rewrite(src, os.path.join(save_path, var), plan, log)
    meta["low_precision"] = {"data_type": data_type.upper(), "plan": {(_node_of(k) or (k,))[0]: v.upper() for k, v in plan.items()}}
    with open(os.path.join(save_path, "saved_model.json"), "w") as f:
        json.dump(meta, f)
    for extra in os.listdir(model_path):                                       # warm-up files etc. travel with the model
        p = os.path.join(model_path, extra)
        if os.path.isfile(p) and extra != "saved_model.json":
            shutil.copy2(p, os.path.join(save_path, extra))
    res["plan"] = plan
    if calib_data:
        from ..serving import Processor
        cfg = {"session_num": 1, "max_batch": max(int(d.shape[0]) for d, _ in calib_data), "model_update_interval_ms": 0}
        a, b = Processor(model_path, cfg, device="cpu"), Processor(save_path, cfg, device="cpu")
        try:
            drift = max(float(abs(a.predict(d, i) - b.predict(d, i)).max()) for d, i in calib_data)
        finally:
            a.close(); b.close()
        res["max_probability_drift"] = drift
        if log:
            log(f"calibration set: max |p_opt - p_fp32| = {drift:.2e} over {len(calib_data)} requests")
        if max_drift is not None and drift > max_drift:
            raise ValueError(f"low-precision drift {drift:.3e} exceeds max_drift {max_drift:.3e}: loosen the plan (opt_config) or the data type")
    return res


def convert_ckpt(ckpt_prefix: str, save_prefix: str, opt_model_path: str, log=None) -> dict:
    """Re-apply the decisions recorded in an optimised model to another bundle of the same model (a later full export's variables, or a delta written
    by ``export_delta*``): "only the parameters changed" updates stay in the optimised format."""
    with open(os.path.join(opt_model_path, "saved_model.json")) as f:
        lp = json.load(f).get("low_precision")
    if not lp:
        raise ValueError(f"{opt_model_path} is not a low-precision model (run optimize() first)")
    r = BundleReader(ckpt_prefix)
    plan = _plan(r.entries, {k: v for k, v in lp["plan"].items() if any((_node_of(n) or (None,))[0] == k for n in r.entries)}, lp["data_type"])
    r.close()
    return _rewrite(ckpt_prefix, save_prefix, plan, log)


# ---- plain checkpoint bundles (training checkpoints) ---------------------------------------------------------------------------------------------
def convert(prefix_in: str, prefix_out: str, dtype: str = "bf16", embeddings_only: bool = False) -> dict:
    r = BundleReader(prefix_in)
    plan = {}
    for name, (dt, shape, nbytes) in r.entries.items():
        n = 1
        for d in shape:
            n *= d
        is_emb = name.endswith("-values") or name.endswith("-sparse_incr_values")
        if dt == "f32" and (is_emb or (not embeddings_only and name.startswith("dense/") and n >= 64)):
            plan[name] = dtype
    r.close()
    if dtype not in ("bf16", "fp16", "int8"):
        raise ValueError("dtype must be bf16 | fp16 | int8")
    return _rewrite(prefix_in, prefix_out, plan, log=None)


def load_tensor(r: BundleReader, name: str) -> torch.Tensor:
    """fp32 view of a (possibly converted) tensor."""
    t = r.read(name)
    if t.dtype == torch.int8 and r.has(name + "/scale"):
        s = r.read(name + "/scale")
        return t.float() * (s.view(-1, *([1] * (t.dim() - 1))) if s.numel() > 1 else s)
    return t.float() if t.dtype in (torch.bfloat16, torch.float16) else t


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="low-precision optimisation of a saved model (directory) or of a checkpoint bundle (prefix)")
    ap.add_argument("--input", required=True); ap.add_argument("--output", required=True)
    ap.add_argument("--data_type", default="bf16", choices=["bf16", "fp16", "int8", "BF16", "FP16", "INT8"]); ap.add_argument("--embeddings_only", action="store_true")
    a = ap.parse_args(argv)
    if os.path.isdir(a.input) and os.path.exists(os.path.join(a.input, "saved_model.json")):
        res = optimize(a.input, a.output, data_type=a.data_type.upper())
        res.pop("plan", None)
        print(res)
    else:
        print(convert(a.input, a.output, a.data_type.lower(), a.embeddings_only))
    return 0


if __name__ == "__main__":
    sys.exit(main())
