"""Difference-quotient Jacobian (UseJacobian 0): lane layout against warp layout (development tool).
usage: gpu_dq_layout.py [name:nsets ...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sbml_odesolver_b200 as so, helpers as H
    name, n = sys.argv[2].split(":"); n = int(n)
    w = H.workload(name, use_jacobian=0)
    m = w["model"]
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)))
    b.set_values(H.workload_values(w, n, seed=77))
    ms = [b.integrate(n) for _ in range(2)]
    st = b.stats(n)
    print(json.dumps(dict(workload=name, n=n, layout=os.environ.get("ODES_WARP", "default"), ms=min(ms), traj_per_s=n / min(ms) * 1e3, failed=int((b.status(n) != 0).sum()),
                          nst=float(st[0].mean()), nje=float(st[3].mean()), info=b.kernel_info())))
else:
    for spec in sys.argv[1:] or ["huang:20000", "mapk:100000"]:
        for warp in ("0", "1"):
            env = dict(os.environ); env["ODES_WARP"] = warp
            r = subprocess.run([sys.executable, __file__, "--one", spec], env=env, capture_output=True, text=True)
            print(r.stdout.strip().split("\n")[-1] if r.returncode == 0 else f"FAILED {spec} warp={warp}: {r.stderr[-300:]}", flush=True)
