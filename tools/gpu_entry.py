"""The reference-shaped one-shot entry (bench.py's e2e_entry leg) alone, for staging-knob sweeps (development tool).
usage: gpu_entry.py [sets]   with ODES_STAGE_THREADS / ODES_STAGE_MB / ODES_STAGE_SLOTS in the environment"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sbml_odesolver_b200 as so, helpers as H, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
w = H.workload("mapk")
w.setdefault("nominal", H.mapk_scan_model()[2])
r = bench.entry_leg(H, so, so.lib(), w, n, 3)
print(json.dumps({"cfg": {k: v for k, v in os.environ.items() if k.startswith("ODES_STAGE")}, "value": r["value"], "d2h_gbs": r["d2h_gbs"],
                  "pinned": r["runhost_pinned_same_shape"]["value"], "ratio": r["ratio_to_runhost_pinned"], "identical": r["identical_to_runhost"]}))
