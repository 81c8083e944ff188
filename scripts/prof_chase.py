"""Band -> tridiagonal chase alone (for ncu): python scripts/prof_chase.py [n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]] + sys.argv[1:]
import scripts.check_sytrd2 as c  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
c.check_stage2(n, vectors=False)
