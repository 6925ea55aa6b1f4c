#!/usr/bin/env bash
# GPU call V4 (one B200): new pipelined-path stress tests (full-GPU grids, ragged sizes 1 .. 4 M points), then the whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-v4}"
( time timeout 900 python -m pytest tests/test_speculate.py -m gpu -q -k "every_sm or ragged" ) > gpurun_out/${T}_pytest_new.log 2>&1
echo "pytest new rc=$?"; tail -40 gpurun_out/${T}_pytest_new.log | cut -c1-300
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
