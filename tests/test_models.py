"""Model-zoo smoke training (cibuild/model-test.sh analogue): every model trains a few steps on CPU and the loss moves."""
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch, taobao_batch
from deeprec_b200.models.zoo import CRITEO_MODELS, TAOBAO_MODELS, build_model
from deeprec_b200.optim import GlobalStep


@pytest.mark.parametrize("name", ["dlrm", "wdl", "deepfm", "dcn", "dcnv2", "masknet"])
def test_criteo_models_train(name):
    torch.manual_seed(0)
    m = build_model(name, device="cpu")
    opt = dr.optim.AdagradOptimizer(m, lr=0.05, global_step=GlobalStep())
    d, ids, y = criteo_batch(256, 13, [50] * 26, seed=1)
    losses = []
    for _ in range(8):
        opt.zero_grad(); l = m.loss(d, ids, y); l.backward(); opt.step(); losses.append(l.item())
    assert all(x == x for x in losses) and losses[-1] < losses[0], losses
    assert all(e.total_count() > 0 for e in dr.optim.collect_embedding_variables(m))


@pytest.mark.parametrize("name", sorted(set(TAOBAO_MODELS)))
def test_taobao_models_train(name):
    torch.manual_seed(0)
    m = build_model(name, device="cpu")
    opt = dr.optim.AdamOptimizer(m, lr=0.01, global_step=GlobalStep())
    b = taobao_batch(128, 10, 500, 800, 20, seed=2)
    losses = []
    for _ in range(8):
        opt.zero_grad(); l = m.loss(b); l.backward(); opt.step(); losses.append(l.item())
    assert all(x == x for x in losses) and losses[-1] < losses[0], losses


def test_train_cli_with_filters_ckpt_and_micro_batch(tmp_path):
    from deeprec_b200.models.train import main
    rc = main(["--model", "wdl", "--steps", "6", "--batch_size", "128", "--device", "cpu", "--ev_filter", "counter", "--ev_elimination", "gstep",
               "--checkpoint", str(tmp_path), "--save_steps", "3", "--micro_batch", "2", "--smartstaged", "--log_every", "0"])
    assert rc == 0
    from deeprec_b200.checkpoint import latest_checkpoint
    assert latest_checkpoint(str(tmp_path)) is not None


def test_mlperf_dlrm_dcn_and_table_variants_train(tmp_path):
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from deeprec_b200.models import train
    # MLPerf DLRM-DCNv2 variant + auto op fusion + work queue over synthetic shards
    assert train.main(["--model", "dlrm_dcn", "--steps", "4", "--batch_size", "64", "--device", "cpu", "--op_fusion", "--workqueue", "--log_every", "0"]) == 0
    # Q-R multi-hash and adaptive embeddings behind the same model code
    assert train.main(["--model", "deepfm", "--steps", "4", "--batch_size", "64", "--device", "cpu", "--multihash", "--log_every", "0"]) == 0
    assert train.main(["--model", "wdl", "--steps", "6", "--batch_size", "64", "--device", "cpu", "--adaptive_emb", "--log_every", "0"]) == 0
    # parquet input (Criteo-shaped) taken through the work queue
    rng = np.random.default_rng(0)
    for part in range(2):
        cols = {"label": rng.integers(0, 2, 256).astype(np.float32)}
        cols.update({f"I{i}": rng.standard_normal(256).astype(np.float32) for i in range(1, 14)})
        cols.update({f"C{i}": rng.integers(0, 500, 256).astype(np.int64) for i in range(1, 27)})
        pq.write_table(pa.table(cols), str(tmp_path / f"part-{part}.parquet"))
    assert train.main(["--model", "dcnv2", "--steps", "6", "--batch_size", "64", "--device", "cpu", "--parquet_dataset", str(tmp_path / "part-*.parquet"),
                       "--workqueue", "--log_every", "0"]) == 0


def test_adaptive_embedding_switches_cold_ids_to_the_ev():
    import deeprec_b200 as dr
    from deeprec_b200.models.zoo import AdaptiveEmbedding
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    ae = AdaptiveEmbedding("adp/C1", 8, hash_bucket_size=16, hot_freq=2, ev_option=None, device=None)
    opt = dr.optim.GradientDescentOptimizer(ae, lr=0.1)
    ids = torch.tensor([3, 19])                          # both hash to bucket 3
    e0 = ae(ids)
    assert torch.equal(e0[0], e0[1]) and torch.equal(e0[0], ae.hashed.weight[3])        # cold: shared static row
    for _ in range(2):
        opt.zero_grad(); ae(ids).sum().backward(); opt.step()
    assert ae.ev.get_frequency(ids).tolist() == [2, 2]
    e1 = ae(ids)                                          # hot now: each id reads its own EmbeddingVariable row
    assert not torch.equal(e1[0], e1[1]) and torch.equal(e1.detach(), ae.ev.table.lookup(ids))


def test_train_cli_evaluates_and_trainer_hooks_can_stop(capsys):
    from deeprec_b200.models import train
    assert train.main(["--model", "esmm", "--steps", "3", "--batch_size", "64", "--device", "cpu", "--log_every", "0", "--eval_steps", "2"]) == 0
    out = capsys.readouterr().out
    assert "Evaluation complete: ACC" in out and "AUC" in out
    assert train.main(["--model", "wdl", "--steps", "2", "--batch_size", "64", "--device", "cpu", "--log_every", "0", "--no_eval"]) == 0
    assert "Evaluation complete" not in capsys.readouterr().out
    # hooks: begin / after_step (stop request) / end
    import deeprec_b200 as dr
    from deeprec_b200.utils import Trainer
    events = []

    class StopAt:
        def begin(self, tr): events.append("begin")
        def after_step(self, tr, step, loss): events.append(step); return step >= 3
        def end(self, tr, step): events.append(("end", step))
    model = torch.nn.Linear(4, 1)
    tr = Trainer(model, dr.optim.GradientDescentOptimizer(model, lr=0.1, global_step=dr.optim.GlobalStep()), lambda m, b: m(b).pow(2).mean(),
                 log_every_n_steps=0, hooks=[StopAt()])
    last = tr.fit((torch.randn(8, 4) for _ in range(100)))
    assert last == 3 and events == ["begin", 1, 2, 3, ("end", 3)]
