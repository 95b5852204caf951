#!/bin/bash
# device timeline of the host-state round trip (VIPRS_B200_E2E_TIMING)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in "VIPRS_B200_NO_ZERO_COPY=1" "X=1"; do
  echo "== $v"
  env $v VIPRS_B200_E2E_TIMING=1 timeout 300 python bench.py --workload c2 --no-extras --no-cpu-baseline --steps 5 2>&1 >/dev/null | grep "viprs_b200 e2e" | tail -8
done
