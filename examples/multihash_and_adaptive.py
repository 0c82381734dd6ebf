"""Compositional (Q-R multi-hash) embeddings and adaptive embeddings."""
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.models.zoo import AdaptiveEmbedding

mh = dr.get_multihash_variable("item", dims=[[64, 8], [64, 8]], complementary_strategy="Q-R", operation="add")
ids = torch.tensor([5, 5 + 64, 5 + 64 * 64])
e = mh.lookup(ids)
print("multi-hash: 4096 ids share", sum(p.numel() for p in mh.parameters()), "parameters; rows differ:", not torch.equal(e[0], e[1]))

ae = AdaptiveEmbedding("query", 8, hash_bucket_size=32, hot_freq=3, ev_option=None, device=None)
opt = dr.optim.AdagradOptimizer(ae, lr=0.1)
x = torch.tensor([1000, 1032])                         # collide in the static table while cold
for step in range(4):
    out = ae(x)
    print(f"step {step}: rows equal (cold, shared bucket) = {torch.equal(out[0], out[1])}, EV frequency {ae.ev.get_frequency(x).tolist()}")
    opt.zero_grad(); out.sum().backward(); opt.step()
assert not torch.equal(ae(x)[0], ae(x)[1])             # hot now: every id owns an EmbeddingVariable row
