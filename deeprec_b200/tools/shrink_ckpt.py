"""``python -m deeprec_b200.tools.shrink_ckpt --input <prefix> --output <prefix>``: drop the filtered (un-admitted) feature
tensors of every EmbeddingVariable from a checkpoint (python/tools/shrink_ckpt_with_filtered_features.py in the reference)."""
from __future__ import annotations

import argparse
import sys

from ..checkpoint.saver import BundleReader, BundleWriter

_FILTERED = ("-keys_filtered", "-freqs_filtered", "-versions_filtered", "-partition_filter_offset")


def shrink(input_prefix: str, output_prefix: str) -> dict:
    r = BundleReader(input_prefix)
    w = BundleWriter(output_prefix)
    kept = dropped = saved = 0
    for name, (_dt, _shape, nbytes) in r.entries.items():
        if any(name.endswith(s) or name.endswith(s + "_incr") for s in _FILTERED):
            dropped += 1; saved += nbytes
            continue
        w.add(name, r.read(name)); kept += 1
    w.close(); r.close()
    return {"kept": kept, "dropped": dropped, "bytes_saved": saved}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True); ap.add_argument("--output", required=True)
    a = ap.parse_args(argv)
    print(shrink(a.input, a.output))
    return 0


if __name__ == "__main__":
    sys.exit(main())
