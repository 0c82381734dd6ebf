"""Run CUDA kernels of this library without a GPU: the SIMT sources are compiled for the host (one host thread per CUDA thread) and
called through the same python wrappers / the same Processor -- what the CPU CI uses to test kernel logic under ASAN / TSAN."""
import os
import tempfile

import numpy as np
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200 import _native
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.zoo import build_model
from deeprec_b200.ops.sparse_ops import sparse_prune_fill, sparse_segment_sum
from deeprec_b200.serving import Processor, export_saved_model_program

sp = dr.SparseIds(torch.tensor([3, -1, 7, 9]), torch.tensor([0, 0, 2, 2]), 4)
want, _ = sparse_prune_fill(sp, default_id=0)                                # torch expressions
with _native.cuda_emulation():
    got, empty = sparse_prune_fill(sp, default_id=0)                          # csrc/cuda/sparse_utils.cu on the host
    seg = sparse_segment_sum(torch.arange(12.).view(4, 3), torch.tensor([0, 1, 3]), torch.tensor([0, 0, 1]), 2)
print("prune + fill-empty-rows:", got.values.tolist(), got.row_ids.tolist(), "== torch:", torch.equal(got.values, want.values))
print("sparse_segment_sum:", seg.tolist())

CARDS = [50] * 26
torch.manual_seed(0)
model = build_model("deepfm", device="cpu", cardinalities=CARDS)
root = tempfile.mkdtemp()
export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root)
d, ids, _ = criteo_batch(16, 13, CARDS, seed=1)
model.eval()
with torch.no_grad():
    ref = torch.sigmoid(model(d, ids)).numpy()
emu = Processor(os.path.join(root, "v1"), {"session_num": 1, "max_batch": 16, "model_update_interval_ms": 0}, device="cuda_emu")
print("DeepFM on the emulated GPU Processor, max |diff| vs module =", float(np.abs(emu.predict(d.numpy(), ids.numpy()) - ref).max()))
emu.close()
