import os, sys, torch
sys.path.insert(0, "/root/repo")
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine
cfg = DLRMConfig(batch_size=65536, overlap_embedding=False)
eng = DLRMEngine(cfg)
for s in range(3):
    d, ids, y = criteo_batch(cfg.batch_size, 13, cfg.cardinalities, seed=s)
    eng.load_batch(d.cuda(), ids.cuda(), y.cuda())
    eng._step_body()
    torch.cuda.synchronize()
    print("step", s, "nuniq", eng.ctx.nuniq.cpu().tolist()[:4], "loss", eng.loss_value(), flush=True)
