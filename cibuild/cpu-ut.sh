#!/usr/bin/env bash
# CPU unit / integration tier (the reference's cibuild/cpu-ut/*.sh): build, lint-free import, every non-GPU test.
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/ -x -q -m "not gpu" "$@"
