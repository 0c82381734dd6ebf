#!/usr/bin/env bash
# Serving tier (cibuild/serving-*.sh): codec, CPU processor, feature stores, HTTP / gRPC front-ends, C client, sanitizer builds.
set -euo pipefail
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests/test_predict_proto.py tests/test_cpu_serving.py tests/test_feature_store.py tests/test_http_serving.py tests/test_examples.py -x -q "$@"
