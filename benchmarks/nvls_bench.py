#!/usr/bin/env python
"""All-reduce micro-benchmark (torchrun): NVLS multimem.ld_reduce kernel vs the one-shot P2P pull vs ncclAllReduce, 64 MB fp32 gradient.
Device-timed (CUDA events), max over ranks; algorithmic bandwidth = bytes / time, bus bandwidth = 2 (W-1)/W * bytes / time."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from deeprec_b200.parallel.p2p import P2PComm
    n = 16 * 1024 * 1024
    res = {}
    for mode in ("nvls", "nvls_2phase", "p2p"):
        os.environ["DEEPREC_NVLS"] = "0" if mode == "p2p" else "1"
        tp = mode == "nvls_2phase"
        comm = P2PComm(rank, world, dev)
        g = comm.alloc_grads(n)
        if mode != "p2p" and comm.nvls is None:
            res[mode] = None
            continue
        torch.manual_seed(rank); g.copy_(torch.randn(n, device=dev))
        out = torch.empty(n, device=dev)
        ref = g.clone(); dist.all_reduce(ref)
        comm.allreduce(out, two_phase=tp); torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        for _ in range(3):
            comm.allreduce(out, two_phase=tp)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            comm.allreduce(out, two_phase=tp)       # (two-phase: in place -- later rounds reduce sums of sums, the timing is what counts)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = {"ms": float(t), "max_err": err}
    x = torch.randn(n, device=dev)
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["nccl"] = {"ms": float(t)}
    if rank == 0:
        byt = n * 4
        for k, v in res.items():
            if v:
                v["algbw_GBs"] = byt / v["ms"] / 1e6; v["busbw_GBs"] = v["algbw_GBs"] * 2 * (world - 1) / world
        print(json.dumps({"bytes": byt, "world": world, **res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
