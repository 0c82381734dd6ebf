"""``python -m deeprec_b200.tools.ckpt_format_transform config.json``: rename tensors of a checkpoint by rewriting ONLY its ``.index`` file -- the
data file (GBs of EmbeddingVariable rows) is not read.  The reference's ``ev_ckpt_transformer`` (tensorflow/tools/embedding_variable/
ckpt_format_transform.cc, README, config.json): same config keys, same use -- bring a checkpoint written under other tensor names (another
framework's dynamic-embedding layout, a refactored model) to the names this model's ``Saver`` / ``CheckpointOption(tensor_name_in_ckpt=...)`` expect.

config.json::

    {"checkpoint_path_prefix": "./ckpt/model.ckpt-1000",
     "output_file": "./ckpt_renamed/model.ckpt-1000.index",
     "tensor_rename_map": {"old/name-keys": "new/name-keys", ...},
     "prefix_rename_map": {"user_emb-1of2": "user_emb/part_0", "user_emb-2of2": "user_emb/part_1"}}      # optional: every tensor starting with the key

``prefix_rename_map`` is the partitioned-table case of the reference's header comment (``a-1of2-keys`` -> ``a/part_0-keys`` ...) without listing each
of a table's tensors.  When ``output_file`` names another prefix than the source, the data file is hard-linked (or symlinked across filesystems) next
to it, so the result is a complete, loadable bundle that still shares the bytes."""
from __future__ import annotations

import json
import os
import sys
from typing import Dict, Optional

_MAGIC = "DEEPREC_B200_BUNDLE"


def transform(checkpoint_path_prefix: str, output_file: str, tensor_rename_map: Optional[Dict[str, str]] = None,
              prefix_rename_map: Optional[Dict[str, str]] = None, link_data: bool = True, log=print) -> dict:
    exact = dict(tensor_rename_map or {})
    prefixes = sorted((prefix_rename_map or {}).items(), key=lambda kv: -len(kv[0]))          # longest prefix wins
    with open(checkpoint_path_prefix + ".index") as f:
        header = f.readline().split()
        if len(header) != 3 or header[0] != _MAGIC:
            raise ValueError(f"{checkpoint_path_prefix}.index is not a checkpoint bundle index")
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    if len(lines) != int(header[2]):
        raise ValueError(f"index lists {header[2]} tensors but has {len(lines)} entries")
    out, seen, renamed, used = [], set(), 0, set()
    for ln in lines:
        name, rest = ln.split("\t", 1)
        new = name
        if name in exact:
            new = exact[name]; used.add(name)
        else:
            for old_p, new_p in prefixes:
                if name.startswith(old_p):
                    new = new_p + name[len(old_p):]; used.add(old_p)
                    break
        if "\t" in new or "\n" in new or not new:
            raise ValueError(f"illegal tensor name {new!r}")
        if new in seen:
            raise ValueError(f"two tensors would be named {new!r}")
        seen.add(new)
        if new != name:
            renamed += 1
            if log:
                log(f"tensor name: {name} -> {new}")
        out.append(new + "\t" + rest)
    missing = [k for k in list(exact) + [p for p, _ in prefixes] if k not in used]
    if missing:
        raise KeyError(f"rename map names tensors / prefixes the checkpoint does not have: {missing}")
    os.makedirs(os.path.dirname(os.path.abspath(output_file)), exist_ok=True)
    tmp = output_file + ".tmp"
    with open(tmp, "w") as f:
        f.write(" ".join(header) + "\n")
        f.write("\n".join(out) + ("\n" if out else ""))
    os.replace(tmp, output_file)
    linked = None
    if link_data and output_file.endswith(".index"):
        dst = output_file[: -len(".index")] + ".data"
        src = os.path.abspath(checkpoint_path_prefix + ".data")
        if os.path.abspath(dst) != src:
            if os.path.lexists(dst):
                os.remove(dst)
            try:
                os.link(src, dst)
            except OSError:
                os.symlink(src, dst)
            linked = dst
    return {"tensors": len(out), "renamed": renamed, "output_file": output_file, "data": linked}


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("Usage: python -m deeprec_b200.tools.ckpt_format_transform config.json", file=sys.stderr)
        return 2
    with open(argv[0]) as f:
        cfg = json.load(f)
    print(transform(cfg["checkpoint_path_prefix"], cfg["output_file"], cfg.get("tensor_rename_map"), cfg.get("prefix_rename_map")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
