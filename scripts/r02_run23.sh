#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02l_pytest.log 2>&1; tail -25 gpurun_out/r02l_pytest.log
