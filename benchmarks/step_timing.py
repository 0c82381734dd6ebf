#!/usr/bin/env python
"""Per-phase CUDA-event timing of eager DLRM engine steps (1 GPU or torchrun): where does a step spend its time, per stream.
  DEEPREC_STEP_TIMING=1 python benchmarks/step_timing.py [--steps 12] [--batch 65536] [--stream fresh|warm]"""
import argparse, json, os, sys
os.environ["DEEPREC_STEP_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12); ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--stream", default="fresh"); ap.add_argument("--prefill", type=int, default=8)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from deeprec_b200.parallel.p2p import P2PComm
        comm = P2PComm(rank, world, dev)
    from deeprec_b200.data.synthetic import criteo_batch
    from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine
    cfg = DLRMConfig(batch_size=a.batch)
    eng = DLRMEngine(cfg, dev, rank, world, comm)
    n = a.prefill + a.steps
    pool = [criteo_batch(cfg.batch_size, 13, cfg.cardinalities, seed=7 + 100003 * rank + (i if a.stream == "fresh" else i % 8)) for i in range(n)]
    for i, (d, ids, y) in enumerate(pool):
        if i == a.prefill:
            eng._events.clear()
        eng.load_batch(d.to(dev), ids.to(dev), y.to(dev)); eng.train_step()
    rep = eng.timing_report(skip=2)
    print(f"[rank {rank}] " + json.dumps({k: round(v, 4) for k, v in rep.items()}), flush=True)
    if world > 1:
        torch.distributed.barrier(); torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
