"""``python -m deeprec_b200.models.train --engine`` for the zoo models (FusedRecEngine behind the modelzoo flag surface), incl. an engine
checkpoint from the CLI.  Written after the round's GPU budget was spent: sorts late on purpose."""
import os

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("model", ["deepfm", "dcn", "din"])
def test_train_cli_engine(model, tmp_path, capsys):
    import deeprec_b200 as dr
    from deeprec_b200.models import train
    dr.embedding_variable.clear_registry()
    ck = str(tmp_path / "eng")
    rc = train.main(["--model", model, "--engine", "--steps", "9", "--batch_size", "512", "--log_every", "4", "--ev_filter", "counter",
                     "--checkpoint", ck, "--save_steps", "4"])
    out = capsys.readouterr().out
    assert rc == 0 and "samples/s" in out and "global_step 8 loss" in out
    assert any(f.startswith("eng-") for f in os.listdir(tmp_path))          # the full checkpoint written at step 4
