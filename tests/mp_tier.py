"""torchrun helper for tests/test_gpu_multi.py: multi-tier tables under model parallelism.  W ranks train two FusedRecEngines on the same batches --
single-tier tables vs a 1024-row-per-rank HBM cache over each rank's own DRAM tier for table 1 (owner-side promotion: k_tier_publish / k_tier_wait /
k_tier_miss_list_mp over peer memory) -- and must see the same losses and end with the same rows on every owner."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import deeprec_b200 as dr
    from deeprec_b200.checkpoint.engine_ckpt import sp_owner
    from deeprec_b200.models.rec_engine import criteo_engine
    from deeprec_b200.models.zoo import build_model
    from deeprec_b200.parallel.p2p import P2PComm
    B, cards = 512, [50, 6000, 7, 300] + [97] * 22
    engs = []
    for tiered in (None, {1: {"cache_rows": 1024, "strategy": 0}}):
        dr.embedding_variable.clear_registry()
        torch.manual_seed(0)
        model = build_model("deepfm", device=dev)
        engs.append(criteo_engine(model, B, table_rows=cards, learning_rate=0.05, tiered=tiered, device=dev, rank=rank, world_size=world,
                                  comm=P2PComm(rank, world, dev)))
    ref, tier = engs
    mgr = tier.tiers[1][0]
    torch.manual_seed(1 + rank)
    batches = []
    for s in range(14):
        lo = (s % 4) * 1500                                    # table 1's working set rotates: 6000 distinct ids >> W * 1024 cache rows
        ids = torch.stack([torch.randint(0, c, (B,), device=dev) for c in cards])
        ids[1] = torch.randint(lo, lo + 1500, (B,), device=dev)
        batches.append((ids, (torch.rand(B, device=dev) < 0.3).float(), {"dense": torch.rand(B, 13, device=dev)}))
    la, lb = [], []
    tier.prefetch(batches[0][0])
    for s, (ids, y, dense) in enumerate(batches):
        for e, l in ((ref, la), (tier, lb)):
            e.load_batch(ids, y, dense)
            e.train_step()
            l.append(e.loss_value())
        if s + 1 < len(batches):
            tier.prefetch(batches[s + 1][0])                   # collective: every rank hands over ITS next batch
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-3, (la, lb)
    st = mgr.stats()
    assert st["demoted_rows"] > 0 and st["promoted_rows"] > 0 and st["evict_passes"] > 0, st
    probe = torch.arange(0, 6000, 7, device=dev)
    mine = sp_owner(probe.cpu(), world).to(dev) == rank
    assert torch.allclose(mgr.lookup(probe)[mine], ref.tables[1].lookup(probe)[mine], atol=1e-4)
    assert tier.tables[1].overflowed() == 0
    dist.barrier()
    if rank == 0:
        print("MP_TIER_OK", st)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
