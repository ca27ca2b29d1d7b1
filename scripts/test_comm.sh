#!/bin/bash
# Communicator smoke test (reference common/comm_core/test.sh + tests/test_comm.py).
here="$(cd "$(dirname "$0")/.." && pwd)"
cd "$here" && exec "${PY:-python}" -m pytest tests/test_comm_api.py -q "$@"
