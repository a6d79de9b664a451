"""Runs one workload's kernel a few times on N sets (for ncu captures).  usage: gpu_small_wl.py name nsets reps"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
name = sys.argv[1]; n = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w = H.workload(name)
b = so.Batch(w["model"], w["opt"], w["vary"], store=list(range(w["model"].neq)))
b.set_values(H.workload_values(w, n, seed=77))
for _ in range(reps):
    print("ms", b.integrate(n))
print("failed", int((b.status(n) != 0).sum()), b.kernel_info())
