"""Records the FP64 FMA peak of this GPU next to the driver's MEASURED_PEAKS.json figures: ODES_measureFP64Peak
(odes_fp64_peak_kernel: independent DFMA chains on every SM), best of several runs.  usage: gpu_fp64_peak.py out.json"""
import datetime, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sbml_odesolver_b200 as so
L = so.lib()
vals = [L.ODES_measureFP64Peak(0, 5) for _ in range(5)]
name = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.max.sm", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
rec = {"fp64_fma_tflops": max(vals), "runs": vals, "gpu": name, "when": datetime.datetime.utcnow().isoformat() + "Z",
       "how": "ODES_measureFP64Peak(device 0, 5 repetitions): odes_fp64_peak_kernel, 8 independent DFMA chains per thread, 2 flop per DFMA, CUDA events; best of 5 calls"}
json.dump(rec, open(sys.argv[1], "w"), indent=1)
print(json.dumps(rec))
