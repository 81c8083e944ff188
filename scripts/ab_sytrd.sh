# A/B timing of the xmca_sytrd variants (XMCA_SYTRD_VARIANT bit mask: 1 = two barriers per column, 2 = transposed slots,
# 4 = release/acquire grid barrier)
set -x
for v in 7 0; do
  XMCA_SYTRD_VARIANT=$v XMCA_PROF_CHECK=1 XMCA_SYTRD_TRACE=1 timeout 300 python scripts/prof_sytrd.py 8192 3 2>&1 | tail -5
done
XMCA_PROF_CHECK=1 timeout 300 python scripts/prof_sytrd.py 16384 2 2>&1 | tail -3
XMCA_PROF_CHECK=1 timeout 300 python scripts/prof_sytrd.py 3000 2 2>&1 | tail -3
XMCA_PROF_CHECK=1 timeout 300 python scripts/prof_sytrd.py 25000 1 2>&1 | tail -3
