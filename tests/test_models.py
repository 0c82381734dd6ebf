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
