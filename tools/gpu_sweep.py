"""Times the MAPK scan kernel for several build-knob settings (development tool).
usage: gpu_sweep.py N "ODES_SETUP_MIN=4 ODES_SETUP_WAIT=2" "ODES_BLOCK=32" ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import sbml_odesolver_b200 as so, helpers as H
    n = int(sys.argv[2])
    m, vary, nominal = H.mapk_scan_model()
    b = so.Batch(m, so.settings(1000.0, 100, 1e-9, 1e-6), vary, store=list(range(m.neq)))
    b.reserve(n)
    b.generate_values(n, 20261019, 0, nominal, -np.ones(len(vary)), np.ones(len(vary)))
    ms = [b.integrate(n) for _ in range(3)]
    st = b.status(n); stats = b.stats(n)
    print(json.dumps(dict(cfg=os.environ.get("ODES_CFG", ""), n=n, ms=min(ms), traj_per_s=n / (min(ms) * 1e-3), failed=int((st != 0).sum()),
                          nst=float(stats[0].mean()), nfe=float(stats[1].mean()), info=b.kernel_info())))
else:
    n = sys.argv[1]
    for cfg in sys.argv[2:] or [""]:
        env = dict(os.environ); env["ODES_CFG"] = cfg
        for kv in cfg.split():
            k, v = kv.split("=", 1); env[k] = v
        r = subprocess.run([sys.executable, __file__, "--one", n], env=env, capture_output=True, text=True)
        print(r.stdout.strip().split("\n")[-1] if r.returncode == 0 else f"FAILED {cfg}: {r.stderr[-400:]}")
