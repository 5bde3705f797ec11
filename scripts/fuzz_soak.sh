#!/bin/bash
# soak run of the three fuzz tests with shifted seeds: scripts/fuzz_soak.sh [first] [last]
cd "$(dirname "$0")/.."
for s in $(seq ${1:-1} ${2:-10}); do
  SOAP_FUZZ_SEED=$s timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k fuzz 2>&1 | tail -4 | sed "s/^/seed $s: /"
done
