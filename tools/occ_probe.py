import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H
w = H.workload("mapk")
b = so.Batch(w["model"], w["opt"], w["vary"], store=list(range(8)))
print(b.kernel_info())
