"""``python -m deeprec_b200.tools.inspect_checkpoint <prefix | directory> [--tensor NAME] [--ev]``: list what a checkpoint bundle
holds (the reference's ``inspect_checkpoint`` + the EmbeddingVariable export format of docs/docs_en/Embedding-Variable-Export-Format.md):
every tensor with dtype / shape / bytes, or one tensor's values, or a per-EmbeddingVariable summary (admitted keys, filtered keys,
dimension, optimizer slot groups, rows per restore partition)."""
from __future__ import annotations

import argparse
import os
import sys
from collections import defaultdict

import torch

from ..checkpoint.saver import BundleReader, latest_checkpoint

_EV_SUFFIXES = ("-keys", "-values", "-freqs", "-versions", "-keys_filtered", "-freqs_filtered", "-versions_filtered", "-partition_offset",
                "-partition_filter_offset")


def ev_summary(r: BundleReader) -> dict:
    """{variable: {"keys", "dim", "filtered_keys", "slots": [...], "bytes"}} -- slot groups are siblings named ``<variable>/<slot>``."""
    groups = defaultdict(dict)
    for name, (dt, shape, nbytes) in r.entries.items():
        for suf in _EV_SUFFIXES:
            if name.endswith(suf):
                groups[name[: -len(suf)]][suf] = (dt, shape, nbytes)
                break
    prim = {g for g in groups if "-keys" in groups[g] and "-values" in groups[g]}
    out = {}
    for g in sorted(prim):
        owner = next((p for p in prim if p != g and g.startswith(p + "/")), None)
        if owner is not None:            # an optimizer slot of another variable
            continue
        shape = groups[g]["-values"][1]
        out[g] = {"keys": int(groups[g]["-keys"][1][0]) if groups[g]["-keys"][1] else 0, "dim": int(shape[1]) if len(shape) > 1 else 1,
                  "filtered_keys": int(groups[g]["-keys_filtered"][1][0]) if "-keys_filtered" in groups[g] and groups[g]["-keys_filtered"][1] else 0,
                  "slots": sorted(s[len(g) + 1:] for s in prim if s.startswith(g + "/")),
                  "bytes": sum(v[2] for s in prim if s == g or s.startswith(g + "/") for v in groups[s].values())}
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint", help="bundle prefix, or a directory (its latest checkpoint is used)")
    ap.add_argument("--tensor", default=None, help="print this tensor")
    ap.add_argument("--ev", action="store_true", help="per-EmbeddingVariable summary")
    ap.add_argument("--partitions", type=int, default=0, help="with --ev: rows each of N restore partitions would load (key %% 1000 %% N)")
    a = ap.parse_args(argv)
    prefix = a.checkpoint
    if os.path.isdir(prefix):
        prefix = latest_checkpoint(prefix) or prefix
    r = BundleReader(prefix)
    try:
        if a.tensor:
            t = r.read(a.tensor)
            torch.set_printoptions(edgeitems=4, linewidth=160)
            print(f"{a.tensor}: dtype={t.dtype} shape={tuple(t.shape)}")
            print(t)
        elif a.ev:
            for name, info in ev_summary(r).items():
                line = f"{name}: keys={info['keys']} dim={info['dim']} filtered_keys={info['filtered_keys']} slots={info['slots']} bytes={info['bytes']}"
                if a.partitions > 1 and info["keys"]:
                    keys = r.read(f"{name}-keys")
                    cnt = torch.bincount(torch.remainder(torch.remainder(keys, 1000), a.partitions), minlength=a.partitions).tolist()
                    line += f" rows_per_partition={cnt}"
                print(line)
        else:
            total = 0
            for name, (dt, shape, nbytes) in sorted(r.entries.items()):
                print(f"{name}\t{dt}\t{list(shape)}\t{nbytes}")
                total += nbytes
            print(f"# {len(r.entries)} tensors, {total} bytes, prefix {prefix}")
    finally:
        r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
