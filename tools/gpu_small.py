"""Runs the MAPK scan kernel a few times on N sets (for ncu captures)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m, vary, nominal = H.mapk_scan_model()
b = so.Batch(m, so.settings(1000.0, 100, 1e-9, 1e-6), vary, store=list(range(m.neq)))
b.reserve(n)
b.generate_values(n, 20261019, 0, nominal, -np.ones(len(vary)), np.ones(len(vary)))
for _ in range(reps):
    print("ms", b.integrate(n))
print("failed", int((b.status(n) != 0).sum()), b.kernel_info())
