#!/usr/bin/env bash
# Smoke-train every modelzoo model for a few steps (the reference's cibuild/model-test.sh).  DEVICE=cuda on a GPU box.
set -euo pipefail
cd "$(dirname "$0")/.."
DEVICE="${DEVICE:-cpu}"
STEPS="${STEPS:-5}"
for m in wdl dlrm deepfm dcn dcnv2 masknet din dien bst dssm esmm mmoe ple dbmtl simple_multitask; do
  echo "== $m"
  python -m deeprec_b200.models.train --model "$m" --steps "$STEPS" --batch_size 256 --device "$DEVICE" --no_eval
done
echo MODEL_TEST_OK
