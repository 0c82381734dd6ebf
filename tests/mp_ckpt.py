"""torchrun helper for tests/test_gpu_multi.py: train a DLRMEngine on W ranks, save a sharded checkpoint, and dump what a restored engine
(of ANY world size) must reproduce: dense parameters and the rows / frequencies of probe keys gathered from their owners."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp_check import gathered_rows  # noqa: E402


def main():
    out_dir = sys.argv[1]
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from deeprec_b200.data.synthetic import criteo_batch
    from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine
    from deeprec_b200.parallel.p2p import P2PComm
    cards = [50, 1000, 7, 300] + [97] * 21 + [200000]
    cfg = DLRMConfig(batch_size=1024, cardinalities=cards, learning_rate=0.05)
    eng = DLRMEngine(cfg, dev, rank, world, P2PComm(rank, world, dev))
    for s in range(3):
        d, ids, y = criteo_batch(cfg.batch_size, 13, cards, seed=100 * rank + s)
        eng.load_batch(d.to(dev), ids.to(dev), y.to(dev)); eng.train_step()
    eng.save(os.path.join(out_dir, "dlrm"))
    probe = torch.arange(0, 300, device=dev)
    rows = {t: tuple(x.cpu() for x in gathered_rows(eng, t, probe, dev)) for t in (0, 1, 3, 25)}
    if rank == 0:
        torch.save({"params": eng.params.cpu(), "rows": rows, "step": eng.global_step()}, os.path.join(out_dir, "expect.pt"))
        print("MP_CKPT_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
