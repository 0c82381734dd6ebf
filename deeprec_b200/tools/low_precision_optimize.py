"""Post-training low-precision conversion of a checkpoint (tools/low_precision_optimize in the reference: BF16 / FP16 / INT8
of SavedModel + checkpoint, including EmbeddingVariables).

``convert(prefix_in, prefix_out, dtype)``: dense tensors and EV ``-values`` are stored as bf16 / fp16 (2x smaller) or
int8 with a per-row (embeddings) / per-tensor (dense) fp32 scale stored next to them as ``<name>/scale``.
``load_tensor(reader, name)`` restores fp32 transparently, so the Saver / serving loaders work on converted bundles."""
from __future__ import annotations

import argparse
import sys

import torch

from ..checkpoint.saver import BundleReader, BundleWriter


def _quant_int8(t: torch.Tensor):
    t = t.float()
    if t.dim() >= 2:
        scale = t.abs().amax(dim=-1, keepdim=True).clamp_min(1e-12) / 127.0
    else:
        scale = (t.abs().max().clamp_min(1e-12) / 127.0).reshape(1)
    q = torch.clamp(torch.round(t / scale), -127, 127).to(torch.int8)
    return q, scale.float()


def convert(prefix_in: str, prefix_out: str, dtype: str = "bf16", embeddings_only: bool = False) -> dict:
    r = BundleReader(prefix_in)
    w = BundleWriter(prefix_out)
    before = after = 0
    for name, (dt, shape, nbytes) in r.entries.items():
        t = r.read(name)
        before += nbytes
        is_emb = name.endswith("-values") or name.endswith("-sparse_incr_values")
        target = is_emb or (not embeddings_only and name.startswith("dense/") and t.dtype == torch.float32 and t.numel() >= 64)
        if t.dtype != torch.float32 or not target:
            w.add(name, t); after += nbytes
            continue
        if dtype in ("bf16", "fp16"):
            q = t.to(torch.bfloat16 if dtype == "bf16" else torch.float16)
            w.add(name, q); after += q.numel() * 2
        elif dtype == "int8":
            q, s = _quant_int8(t)
            w.add(name, q); w.add(name + "/scale", s); after += q.numel() + s.numel() * 4
        else:
            raise ValueError("dtype must be bf16 | fp16 | int8")
    w.close(); r.close()
    return {"bytes_before": before, "bytes_after": after, "ratio": after / max(1, before)}


def load_tensor(r: BundleReader, name: str) -> torch.Tensor:
    """fp32 view of a (possibly converted) tensor."""
    t = r.read(name)
    if t.dtype == torch.int8 and r.has(name + "/scale"):
        return t.float() * r.read(name + "/scale")
    return t.float() if t.dtype in (torch.bfloat16, torch.float16) else t


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True); ap.add_argument("--output", required=True)
    ap.add_argument("--data_type", default="bf16", choices=["bf16", "fp16", "int8"]); ap.add_argument("--embeddings_only", action="store_true")
    a = ap.parse_args(argv)
    print(convert(a.input, a.output, a.data_type, a.embeddings_only))
    return 0


if __name__ == "__main__":
    sys.exit(main())
