#!/bin/bash
# A/B timing of the xmca_sytrd tuning knobs on one GPU (used for the measurements quoted in DESIGN.md §6):
#   XMCA_SYTRD_VARIANT   bit mask: 1 = two grid barriers per column, 2 = transposed partial-sum slots,
#                        4 = release/acquire counter barrier (default 7; 0 = the three-barrier kernel)
#   XMCA_SYTRD_TILE_MIN  tile-major passes while the trailing size exceeds this (default 4096)
#   XMCA_SYTRD_KEEP_MB   plain instead of evict-first loads once the trailing matrix is at most this large (default 88)
#   XMCA_PROF_PAIR=1     also time the batched two-problem call
# usage: bash scripts/ab_sytrd.sh [n]
n=${1:-8192}
for v in 0 7; do
  echo "== XMCA_SYTRD_VARIANT=$v"
  XMCA_SYTRD_VARIANT=$v XMCA_PROF_CHECK=1 XMCA_SYTRD_TRACE=1 timeout 300 python scripts/prof_sytrd.py $n 3 2>&1 | tail -3
done
echo "== batched pair"
XMCA_PROF_PAIR=1 XMCA_PROF_CHECK=1 timeout 300 python scripts/prof_sytrd.py $n 2 2>&1 | tail -2
