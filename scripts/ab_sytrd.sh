# A/B timing of xmca_sytrd tuning knobs: XMCA_SYTRD_TILE_MIN (tile-major while the trailing size exceeds it),
# XMCA_SYTRD_KEEP_MB (plain instead of evict-first loads once the trailing matrix is at most this large)
set -x
for cfg in "4096 0" "4096 88" "2048 88" "1024 88" "2048 0" "512 88" "1024 110"; do
  set -- $cfg
  XMCA_SYTRD_TILE_MIN=$1 XMCA_SYTRD_KEEP_MB=$2 XMCA_PROF_CHECK=1 timeout 300 python scripts/prof_sytrd.py 8192 3 2>&1 | tail -2
done
