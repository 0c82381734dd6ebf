"""Synchronous training on 2 processes (gloo on CPU, nccl on GPUs): embeddings model-parallel (table t lives on rank t % world),
dense net data-parallel with a bucketed gradient all-reduce.  Run directly: the script launches its own workers."""
import os
import sys

import torch

import _path  # noqa: F401  (repository root on sys.path)


def worker():
    import deeprec_b200 as dr
    from deeprec_b200.parallel import CollectiveStrategy
    st = CollectiveStrategy()
    torch.manual_seed(0)
    evs = [dr.get_embedding_variable(f"C{t}", 8, seed=t) for t in range(4)]
    dense = torch.nn.Sequential(torch.nn.Linear(32, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    st.broadcast_parameters(dense)
    opt = dr.optim.AdagradOptimizer(dense.parameters(), evs, lr=0.05)
    g = torch.Generator().manual_seed(10 + st.rank)                    # every rank reads its own shard of the data
    for step in range(10):
        ids = [torch.randint(0, 100, (64,), generator=g) for _ in evs]
        with st.scope(), st.embedding_scope():
            embs = dr.group_embedding_lookup_sparse(evs, [dr.SparseIds.from_dense(i) for i in ids], ["sum"] * 4)
        loss = (dense(torch.cat(embs, 1)).squeeze(-1) - 1.0).pow(2).mean()
        opt.zero_grad(); loss.backward()
        st.allreduce_gradients(list(dense.parameters()), average=True)
        opt.step()
    owned = [t for t, e in enumerate(evs) if e.total_count() > 0]
    print(f"rank {st.rank}/{st.world_size}: loss {loss.item():.4f}, owns tables {owned}", flush=True)
    assert owned == [t for t in range(4) if t % st.world_size == st.rank]


if __name__ == "__main__":
    if "RANK" in os.environ:
        worker()
    else:
        from deeprec_b200.parallel import launch
        sys.exit(launch.main(["--nproc", "2", os.path.abspath(__file__)]))
