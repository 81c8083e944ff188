# A/B timing of the batched tridiagonalisation: phase offset between the two groups (clocks)
set -x
for dl in 0 15000 30000 45000 60000; do
  XMCA_SYTRD_BATCH_DELAY=$dl XMCA_PROF_PAIR=1 timeout 300 python scripts/prof_sytrd.py 8192 2 2>&1 | tail -1
done
