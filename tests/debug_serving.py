import os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine
from deeprec_b200.data import criteo_batch
from deeprec_b200.serving import Processor, export_saved_model
cards = [50, 1000, 7, 300] + [97] * 22
eng = DLRMEngine(DLRMConfig(batch_size=512, cardinalities=cards, learning_rate=0.05))
for s in range(4):
    d, ids, y = criteo_batch(512, 13, cards, seed=s)
    eng.load_batch(d.cuda(), ids.cuda(), y.cuda()); eng.train_step()
root = tempfile.mkdtemp()
export_saved_model(eng, os.path.join(root, "v1"), version=4, root=root)
proc = Processor(os.path.join(root, "v1"), {"session_num": 1, "max_batch": 256, "model_update_interval_ms": 0})
eng.load_batch(d.cuda(), ids.cuda(), torch.zeros(512, device="cuda"))
ref = eng.predict().cpu().numpy().copy()
got = proc.predict(d.numpy(), ids.numpy())
bad = np.isnan(got)
print("nan count", bad.sum(), "first bad", np.nonzero(bad)[0][:10], "maxdiff(ok)", np.abs(got[~bad] - ref[~bad]).max() if (~bad).any() else None)
got2 = proc.predict(d.numpy()[:256], ids.numpy()[:, :256])
print("B=256 nan", np.isnan(got2).sum(), "maxdiff", np.nanmax(np.abs(got2 - ref[:256])))
got3 = proc.predict(d.numpy()[:128], ids.numpy()[:, :128])
print("B=128 nan", np.isnan(got3).sum())
