#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
( VIPRS_B200_TILE_LIMIT=255 VIPRS_B200_TILE_ROWS=256 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q 2>&1 ) | tail -40
