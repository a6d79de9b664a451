"""GPU tier (run with -m gpu on the B200 box): the CUDA kernel, called through the C ABI of libodes_b200.so,
against the oracle on the same seeded inputs.

Gate (north star): every output point within max(10 rtol |x|, 10 atol) of the reference's own CVODE integration,
identical event firing order, identical failure flags.  The kernel replays CVODE's arithmetic exactly (IEEE
operations in the reference's order, no FMA contraction, glibc's pow restated), so the tests demand more than the
gate: the same per-trajectory work counters (nst, nfe, nsetups, nje, nni, ncfn, netf) and BIT-IDENTICAL stored
values against the oracle driven with the same emitted right-hand side (the reference's compiled mode)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
import sbml_odesolver_b200 as so

pytestmark = pytest.mark.gpu


def _gpu_run(w, vals, store=None):
    b = so.Batch(w["model"], w["opt"], w["vary"], store=store)
    n = b.set_values(vals)
    b.integrate(n)
    r = dict(out=b.results(n), status=b.status(n), stats=b.stats(n).T, fired=b.event_log(n).T, info=b.kernel_info())
    b.close()
    return r


# The checker of the GPU tier is the reference's OWN arithmetic: the vendored cvode.c / cvdense.c / smalldense.c compiled
# unmodified into oracle/_ref/libcvode_ref.so (backend REF).  The restated port (equal to it bit for bit in the CPU tier,
# tests/test_oracle_golden.py) is only the fallback where oracle/_ref was not built.
BACKEND = H.gpu_tier_backend()


def _oracle(w, vals, **kw):
    m = w["model"]
    return H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=BACKEND,
                        compiled_so=H.compile_c_model(m, w["name"]), threads=8, **kw)


def test_checker_is_the_vendored_cvode():
    """The GPU box gets oracle/_ref from the snapshot: every BDF parity test below runs against the reference's CVODE."""
    assert H.have_ref() and BACKEND == H.REF


# repressilator at the BASELINE shape (configs[2]): t = 0..1000, 1000 output points
@pytest.mark.parametrize("name,nsets,nout", [("mapk", 1024, None), ("repressilator", 96, 200), ("repressilator", 64, 1000),
                                             ("goodwin", 256, None), ("huang", 128, None)])
def test_parity_with_oracle(name, nsets, nout):
    w = H.workload(name, nout=nout)
    vals = H.workload_values(w, nsets, seed=2026)
    got = _gpu_run(w, vals)
    ref = _oracle(w, vals)
    want = np.transpose(ref["out"], (1, 2, 0))
    assert np.array_equal(got["status"], ref["status"])
    rep = H.parity_report(got["out"], want, got["stats"], ref["stats"], w["rtol"], w["atol"])
    bitident = float(np.mean([np.array_equal(got["out"][:, :, k], want[:, :, k], equal_nan=True) for k in range(nsets)]))
    print(name, {k: v for k, v in rep.items() if k not in ("ratio", "same_path")}, "bit-identical sets", bitident, got["info"])
    # tolerance gate: max(10 rtol |x|, 10 atol) at every output point -- and in fact bit for bit
    assert rep["frac_within"] == 1.0 and rep["max_ratio"] <= 1.0, rep
    assert rep["frac_same_path"] == 1.0, rep
    assert bitident == 1.0, (bitident, rep)


@pytest.mark.parametrize("name,nsets,nout,layout", [("mapk", 512, None, "lane"), ("repressilator", 64, 200, "lane"), ("huang", 64, None, "lane"),
                                                    ("huang", 64, None, "warp"), ("mapk", 256, None, "warp")])
def test_difference_quotient_jacobian(monkeypatch, name, nsets, nout, layout):
    """UseJacobian 0: CVDENSE's CVDenseDQJac (cvdense.c:479-537, src/cvodeSolver.c:620-623) in the lane kernel and in the warp kernel (the default from 19 equations on), against the vendored CVODE with its own difference quotients."""
    monkeypatch.setenv("ODES_WARP", "1" if layout == "warp" else "0")      # default: warp from 19 equations on, as with a Jacobian
    w = H.workload(name, nout=nout, use_jacobian=0)
    vals = H.workload_values(w, nsets, seed=31)
    got = _gpu_run(w, vals)
    ref = _oracle(w, vals)
    assert np.array_equal(got["status"], ref["status"]) and (ref["status"] == 0).all()
    assert np.array_equal(got["stats"], ref["stats"]) and (ref["stats"][:, 3] > 0).all()
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    print(name, layout, got["info"])


def test_entire_headline_batch_is_bit_identical_to_vendored_cvode():
    """BASELINE configs[1] at its full size: all 10^6 trajectories of the headline batch -- every stored value of every output
    row, the status flags and CVODE's work counters -- against the vendored CVODE run on the host threads over the same sets
    in slabs of 10^5 sets (about 30 s on 16 cores; 6.5 GB for the device's result block, 3 GB per oracle slab)."""
    if not H.have_ref():
        pytest.skip("oracle/_ref not built")
    with open("/proc/meminfo") as f:
        avail_gb = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] / 1048576
    if avail_gb < 24:
        pytest.skip("needs 24 GB of host memory")
    n = 1_000_000
    m, vary, nominal = H.mapk_scan_model()
    opt = so.settings(1000.0, 100, 1e-9, 1e-6)
    b = so.Batch(m, opt, vary, store=list(range(m.neq)))
    b.reserve(n)
    b.generate_values(n, 20261019, 0, nominal, -np.ones(len(vary)), np.ones(len(vary)))
    b.integrate(n)
    out, status, stats = b.results(n), b.status(n), b.stats(n)
    vals = np.ascontiguousarray(b.get_values(n))
    b.close()
    assert int((status != 0).sum()) == 0
    cso = H.compile_c_model(m, "mapk_scan")
    for k in range(0, n, 100_000):
        ref = H.oracle_run(m, opt, vary, vals[k:k + 100_000], 100, backend=H.REF, compiled_so=cso, threads=os.cpu_count() or 1)
        assert np.array_equal(status[k:k + 100_000], ref["status"])
        assert np.array_equal(stats[:, k:k + 100_000].T, ref["stats"])
        assert np.array_equal(out[:, :, k:k + 100_000], np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0))), k


def test_event_order_identical():
    w = H.workload("events")
    vals = H.workload_values(w, 2048, seed=9)
    got = _gpu_run(w, vals)
    m = w["model"]
    ref = _oracle(w, vals, want_margin=True)
    # sets whose trigger expression comes within 100*max(rtol |x|, atol) of its threshold at an output point are
    # ill-conditioned and reported separately
    safe = (ref["margin"] > 100 * np.maximum(w["rtol"] * 1.0, w["atol"])).all(axis=1)
    assert safe.mean() > 0.9
    assert np.array_equal(got["fired"][safe], ref["fired"][safe])
    mism = int((got["fired"][~safe] != ref["fired"][~safe]).any(axis=1).sum())
    print("event parity: sets", len(vals), "ill-conditioned", int((~safe).sum()), "of which mismatching", mism)
    want = np.transpose(ref["out"], (1, 2, 0))
    assert H.parity(got["out"][:, :, safe], want[:, :, safe], w["rtol"], w["atol"]) <= 1.0
    # nominal set: the firing sequence of SURVEY appendix D
    nom = _gpu_run(w, np.array([[1.0, 0.0]]))
    f = nom["fired"][0]
    assert [(e, i) for i in range(101) for e in (0, 1) if f[i] >> e & 1] == \
        [(1, 7), (0, 24), (1, 25), (1, 34), (0, 48), (1, 51), (1, 63), (0, 72), (1, 77), (1, 95), (0, 96)]


def test_golden_testrun_through_the_kernel():
    """examples/testrun.out:28-129 (atol 1e-9, rtol 1e-4): the kernel prints the same 6 digits as the reference."""
    import os
    gold = np.loadtxt(os.path.join(H.ROOT, "tests", "golden", "mapk_testrun.txt"), comments="#")
    m = so.Model(H.model_path("mapk_cascade.xml"))
    opt = so.settings(1000.0, 100, 1e-9, 1e-4, mxstep=1000)
    b = so.Batch(m, opt, [], store=list(range(8)))
    b.reserve(32)
    b.integrate(1)
    out = b.results(1)[:, :, 0]
    got = np.array([[float("%g" % v) for v in row] for row in out])
    assert np.array_equal(got, gold[:, 1:])
    b.close()


def test_ragged_and_edge_batches():
    w = H.workload("mapk", nout=10)
    m = w["model"]
    vals = H.workload_values(w, 97, seed=4)                      # not a multiple of the warp or block size
    ref = _oracle(w, vals)
    want = np.transpose(ref["out"], (1, 2, 0))
    b = so.Batch(m, w["opt"], w["vary"])
    for n in (1, 31, 33, 97):
        b.set_values(vals[:n])
        b.integrate(n)
        assert H.parity(b.results(n), want[:, :, :n], w["rtol"], w["atol"]) <= 1.0
        assert (b.status(n) == 0).all()
    b.integrate(0)                                               # empty batch is a no-op
    b.close()


def test_failure_flags_and_nan_rows():
    w = H.workload("mapk", nout=4)
    vals = H.workload_values(w, 40, seed=5)
    opt = so.settings(w["end"], 4, w["atol"], w["rtol"], mxstep=15)
    b = so.Batch(w["model"], opt, w["vary"])
    b.set_values(vals)
    b.integrate(40)
    st, out = b.status(40), b.results(40)
    ref = H.oracle_run(w["model"], opt, w["vary"], vals, 4, backend=H.gpu_tier_backend())
    assert np.array_equal(st, ref["status"]) and (st == -4).all()
    assert np.array_equal(np.isnan(out), np.isnan(np.transpose(ref["out"], (1, 2, 0))))
    b.close()


def test_run_host_pipeline_and_one_shot_entry():
    """ODESBatch_runHost (chunked H2D / kernel / D2H overlap) and IntegratorInstance_integrateBatch give the plan's results."""
    w = H.workload("mapk", nout=20)
    m, L = w["model"], so.lib()
    vals = H.workload_values(w, 1000, seed=6)
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(8)))
    b.set_values(vals)
    b.integrate(1000)
    direct = b.results(1000)
    res = np.zeros((1000, direct.shape[0], direct.shape[1]))          # [set][row][value], the library's layout
    st = np.zeros(1000, dtype=np.int32)
    # chunk > 0: launches of that many sets, finished groups of sets stream back while the kernel runs (group size
    # 2^ODES_GROUP_SHIFT, ragged last group); chunk < 0: the double-buffered scheme with one copy per launch
    for chunk, shift in ((256, "5"), (0, "4"), (0, "14"), (992, "6"), (-256, "14")):
        monkey_env = dict(ODES_GROUP_SHIFT=shift)
        os.environ.update(monkey_env)
        res[:] = -1.0
        st[:] = -1
        b.run_host(vals, res, st, chunk=chunk)
        assert np.array_equal(np.transpose(res, (1, 2, 0)), direct) and (st == 0).all(), (chunk, shift)
    os.environ.pop("ODES_GROUP_SHIFT")
    b.close()
    ii = L.IntegratorInstance_create(m.om, w["opt"])
    assert ii
    vis = (C.c_void_p * len(w["vary"]))(*[L.ODEModel_getVariableIndexByNum(m.om, i) for i in w["vary"]])
    out = np.zeros((64, 21, m.nvalues))
    st = np.zeros(64, dtype=np.int32)
    failed = L.IntegratorInstance_integrateBatch(ii, 64, len(w["vary"]), vis, vals[:64].ctypes.data, out.ctypes.data, st.ctypes.data)
    assert failed == 0 and (st == 0).all()
    assert np.array_equal(out[:, :, :8], np.transpose(direct[:, :, :64], (2, 0, 1)))
    ref = H.oracle_run(m, w["opt"], w["vary"], vals[:64], 20, backend=H.gpu_tier_backend())
    assert H.parity(np.transpose(out, (1, 2, 0)), np.transpose(ref["out"], (1, 2, 0)), w["rtol"], w["atol"]) <= 1.0
    L.IntegratorInstance_free(ii)


def test_one_shot_entry_keeps_its_plans_on_the_model():
    """IntegratorInstance_integrateBatch, the reference-shaped entry (caller's malloc'd arrays, all model values): the second
    call with the same model / settings / varied set compiles nothing, loads no module and allocates nothing on the device
    (buffers cached on ii / om); changed tolerances and a changed time grid reuse the plan and still give the
    vendored CVODE's bits."""
    L = so.lib()
    w = H.workload("mapk", nout=20)
    m = w["model"]
    vals = H.workload_values(w, 96, seed=77)
    ii = L.IntegratorInstance_create(m.om, w["opt"])
    vis = (C.c_void_p * len(w["vary"]))(*[L.ODEModel_getVariableIndexByNum(m.om, i) for i in w["vary"]])

    def call(opt_ii):
        out = np.full((96, 21, m.nvalues), -1.0)
        st = np.full(96, -99, dtype=np.int32)
        assert L.IntegratorInstance_integrateBatch(opt_ii, 96, len(w["vary"]), vis, vals.ctypes.data, out.ctypes.data, st.ctypes.data) == 0
        return out
    first = call(ii)
    c1 = so.counters()
    second = call(ii)
    c2 = so.counters()
    assert np.array_equal(first, second)
    for k in ("nvrtc_compiles", "cubin_cache_loads", "module_loads", "device_allocations", "plans_created"):
        assert c2[k] == c1[k], (k, c1, c2)
    assert c2["plan_cache_hits"] == c1["plan_cache_hits"] + 1 and c2["kernel_launches"] > c1["kernel_launches"]
    ref = _oracle(w, vals)
    assert np.array_equal(first, ref["out"], equal_nan=True)
    # tighter tolerances and another end time through the same cached plan
    opt2 = so.settings(w["end"] / 2, 20, 1e-10, 1e-7)
    ii2 = L.IntegratorInstance_create(m.om, opt2)
    third = call(ii2)
    c3 = so.counters()
    assert c3["plans_created"] == c2["plans_created"] and c3["nvrtc_compiles"] == c2["nvrtc_compiles"] and c3["module_loads"] == c2["module_loads"]
    w2 = dict(w, opt=opt2)
    assert np.array_equal(third, _oracle(w2, vals)["out"], equal_nan=True)
    L.ODEModel_freeKernelCache(m.om)
    fourth = call(ii)
    assert np.array_equal(fourth, first) and so.counters()["plans_created"] == c3["plans_created"] + 1
    L.IntegratorInstance_free(ii2)
    L.IntegratorInstance_free(ii)


def test_full_size_properties():
    """BASELINE config 2 at a size the oracle cannot cover: size-independent properties of the model.
    MAPK conserves MKKK+MKKK_P, MKK+MKK_P+MKK_PP and MAPK+MAPK_P+MAPK_PP (stoichiometry), concentrations stay
    non-negative up to the tolerance, every set finishes, and a re-run is bit-identical (deterministic kernel)."""
    w = H.workload("mapk")
    n = 200_000
    b = so.Batch(w["model"], w["opt"], w["vary"], store=list(range(8)))
    b.reserve(n)
    nv = len(w["vary"])
    b.generate_values(n, 20261019, 0, w["nominal"], -np.ones(nv), np.ones(nv))
    b.integrate(n)
    out, st = b.results(n), b.status(n)
    assert (st == 0).all() and not np.isnan(out).any()
    for grp, total in (((0, 1), 100.0), ((2, 3, 4), 300.0), ((5, 6, 7), 300.0)):
        s = sum(out[:, k, :] for k in grp)
        assert np.abs(s - total).max() <= 10 * 1e-6 * total
    assert out.min() > -10 * 1e-9 - 1e-5
    # a shard regenerates its slice: sets [1000, 1256) of the big batch == a batch generated with firstSet = 1000
    b2 = so.Batch(w["model"], w["opt"], w["vary"], store=list(range(8)))
    b2.reserve(256)
    b2.generate_values(256, 20261019, 1000, w["nominal"], -np.ones(nv), np.ones(nv))
    b2.integrate(256)
    assert np.array_equal(b2.results(256), out[:, :, 1000:1256])
    b.integrate(n)
    assert np.array_equal(b.results(n), out)
    b.close(); b2.close()


def test_feature_model_on_gpu(tmp_path):
    """Function definitions, initial assignment, rate + piecewise assignment rules, stoichiometryMath, a Hill-type
    power, a logical trigger and an event that restarts the solver: GPU vs oracle, bit for bit."""
    from test_frontend_models import FEATURES
    p = tmp_path / "features.xml"
    p.write_text(FEATURES)
    m = so.Model(str(p))
    vary = [m.index(n) for n in ("vmax", "km", "kd", "A")]
    rng = np.random.default_rng(3)
    n = 4096
    vals = np.stack([rng.uniform(0.5, 3, n), rng.uniform(0.2, 2, n), rng.uniform(0.05, 1, n), rng.uniform(1, 6, n)], axis=1)
    opt = so.settings(40.0, 80, 1e-9, 1e-6)
    b = so.Batch(m, opt, vary)
    b.set_values(vals)
    b.integrate(n)
    out, st, stats, fired = b.results(n), b.status(n), b.stats(n).T, b.event_log(n).T
    ref = H.oracle_run(m, opt, vary, vals, 80, compiled_so=H.compile_c_model(m, "features"), threads=8, backend=H.gpu_tier_backend())
    assert np.array_equal(st, ref["status"]) and np.array_equal(stats, ref["stats"])
    assert np.array_equal(fired, ref["fired"]) and ref["fired"].any()
    assert np.array_equal(out, np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    b.close()


@pytest.mark.parametrize("n", [3, 12, 17, 20, 32, 40])
def test_chain_models_on_gpu(tmp_path, n):
    """n = 3 and 12 use tensor memory with one and two chunks per matrix column (12: all 512 columns of one block per
    SM), n = 17 falls back to the global scratch, n = 20 and 32 take the warp-per-trajectory layout, n = 40 (beyond a warp) the lane layout with the global scratch: all
    bit-identical to the oracle."""
    from test_frontend_models import chain_model_xml
    p = tmp_path / f"chain{n}.xml"
    p.write_text(chain_model_xml(n))
    m = so.Model(str(p))
    vary = [m.index(f"v{i}") for i in range(1, n)] + [m.index("K")]
    rng = np.random.default_rng(n)
    ns = 600
    vals = m.base_values()[vary][None, :] * np.exp2(rng.uniform(-1, 1, size=(ns, len(vary))))
    opt = so.settings(30.0, 30, 1e-9, 1e-6)
    b = so.Batch(m, opt, vary)
    b.set_values(vals)
    b.integrate(ns)
    out, st, stats = b.results(ns), b.status(ns), b.stats(ns).T
    ref = H.oracle_run(m, opt, vary, vals, 30, compiled_so=H.compile_c_model(m, f"chain{n}"), threads=8, backend=H.gpu_tier_backend())
    print(n, b.kernel_info())
    assert (st == 0).all() and np.array_equal(stats, ref["stats"])
    assert np.array_equal(out, np.transpose(ref["out"], (1, 2, 0)))
    b.close()


@pytest.mark.parametrize("n", [4, 8, 10])
def test_row_exchanges_in_the_tensor_memory_lu(tmp_path, n):
    """Models whose Newton matrices need partial pivoting (product stoichiometry 3): the left-looking LU of the lane kernel
    exchanges rows of the multiplier columns it keeps in tensor memory; bit-identical to the vendored CVODE."""
    from test_frontend_models import amplifier_case
    m, vary, vals, opt = amplifier_case(tmp_path, n, 1500)
    b = so.Batch(m, opt, vary)
    b.set_values(vals)
    b.integrate(len(vals))
    out, st, stats = b.results(len(vals)), b.status(len(vals)), b.stats(len(vals)).T
    ref = H.oracle_run(m, opt, vary, vals, 20, compiled_so=H.compile_c_model(m, f"amp{n}"), threads=8, backend=H.gpu_tier_backend())
    print(n, b.kernel_info())
    assert (st == 0).all() and np.array_equal(stats, ref["stats"])
    assert np.array_equal(out, np.transpose(ref["out"], (1, 2, 0)))
    b.close()


@pytest.mark.parametrize("name,nsets,nout", [("mapk", 300, None), ("goodwin", 200, None), ("events", 400, None), ("repressilator", 64, 200)])
def test_warp_layout_forced_on_small_models(monkeypatch, name, nsets, nout):
    """bdf_warp_kernel.cuh (one trajectory per warp; automatic for n > 12) on the small models: solver restarts after
    events, pow in rate laws, failure flags -- bit-identical to the oracle like the lane layout."""
    monkeypatch.setenv("ODES_WARP", "1")
    w = H.workload(name, nout=nout)
    vals = H.workload_values(w, nsets, seed=515)
    got = _gpu_run(w, vals)
    assert got["info"]["threads_per_block"] == 128 and got["info"]["smem_bytes_per_block"] > 0
    ref = _oracle(w, vals)
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["stats"], ref["stats"])
    if name == "events":
        assert np.array_equal(got["fired"], ref["fired"]) and ref["fired"].any()
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)


def test_warp_layout_failure_flags(monkeypatch):
    monkeypatch.setenv("ODES_WARP", "1")
    w = H.workload("mapk", nout=4)
    m = w["model"]
    opt = so.settings(1000.0, 4, 1e-9, 1e-6, mxstep=15)
    vals = H.workload_values(w, 40, seed=3)
    b = so.Batch(m, opt, w["vary"])
    b.set_values(vals)
    b.integrate(40)
    out, st = b.results(40), b.status(40)
    ref = H.oracle_run(m, opt, w["vary"], vals, 4, compiled_so=H.compile_c_model(m, "mapk"), backend=H.gpu_tier_backend())
    assert (st == -4).all() and np.array_equal(st, ref["status"])
    assert np.array_equal(out, np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    b.close()


def test_run_host_streams_results_of_the_warp_layout():
    """The warp-per-trajectory kernel raises the same completion flags (huang96, failures included: mxstep too small for some sets)."""
    w = H.workload("huang", nout=10)
    m = w["model"]
    opt = so.settings(10.0, 10, 1e-9, 1e-6, mxstep=40)
    vals = H.workload_values(w, 700, seed=8)
    b = so.Batch(m, opt, w["vary"], store=list(range(m.neq)))
    b.set_values(vals)
    b.integrate(700)
    direct, dst = b.results(700), b.status(700)
    res = np.zeros((700, direct.shape[0], direct.shape[1]))
    st = np.zeros(700, dtype=np.int32)
    os.environ["ODES_GROUP_SHIFT"] = "5"
    b.run_host(vals, res, st, chunk=0)
    os.environ.pop("ODES_GROUP_SHIFT")
    assert np.array_equal(np.transpose(res, (1, 2, 0)), direct, equal_nan=True) and np.array_equal(st, dst)
    b.close()


@pytest.mark.parametrize("warp", [0, 1])
def test_steady_state_halt_on_gpu(monkeypatch, warp):
    """Per-trajectory steady-state halting (CvodeSettings_setHaltOnSteadyState): same last row as the oracle, rows after
    it not-a-number, status 0 -- in both layouts."""
    monkeypatch.setenv("ODES_WARP", str(warp))
    w = H.workload("huang", nout=40)
    m = w["model"]
    opt = so.settings(400.0, 40, 1e-9, 1e-6, steady_state=1, ss_threshold=1e-5)
    vals = H.workload_values(w, 300, seed=12)
    b = so.Batch(m, opt, w["vary"])
    b.set_values(vals)
    b.integrate(300)
    out, st = b.results(300), b.status(300)
    b.close()
    ref = H.oracle_run(m, opt, w["vary"], vals, 40, compiled_so=H.compile_c_model(m, "huang"), threads=8, backend=H.gpu_tier_backend())
    want = np.transpose(ref["out"], (1, 2, 0))
    assert (st == 0).all() and np.isnan(want[-1, 0, :]).mean() > 0.5        # most sets settle inside the grid
    assert np.array_equal(out, want, equal_nan=True)


def test_single_instance_steady_state_matches_the_batch():
    """IntegratorInstance_integrate with HaltOnSteadyState on a model with NEGATIVE rates (huang96): the host's
    IntegratorInstance_checkSteadyState (src/integratorInstance.c:1442-1507: signed rates against the mean of the |rates|)
    must stop at the output time at which the device's steady_f and the oracle stop -- a test on |rate| - mean would stop
    earlier."""
    L = so.lib()
    w = H.workload("huang", nout=40)
    m = w["model"]
    opt = so.settings(400.0, 40, 1e-9, 1e-6, steady_state=1, ss_threshold=1e-5)
    vals = w["nominal"][None, :].copy()
    ref = H.oracle_run(m, opt, w["vary"], vals, 40, compiled_so=H.compile_c_model(m, "huang"), backend=H.gpu_tier_backend())
    rows = np.where(~np.isnan(ref["out"][0, :, 0]))[0]
    assert 0 < rows[-1] < 40, "the nominal huang96 run must settle inside the grid for this test"
    t_halt = H.tpoints(opt, 40)[rows[-1]]
    ii = L.IntegratorInstance_create(m.om, opt)
    assert ii
    L.IntegratorInstance_integrate(ii)
    assert L.IntegratorInstance_getTime(ii) == t_halt, (L.IntegratorInstance_getTime(ii), t_halt)
    assert "120501" in so.errors() and "Steady state found" in so.errors()     # SOLVER_MESSAGE_STEADYSTATE_FOUND
    L.SolverError_clear()
    L.IntegratorInstance_free(ii)


@pytest.mark.parametrize("name,n,nout", [("mapk", 6000, 10), ("huang", 640, 20), ("goodwin", 2000, 20)])
def test_all_devices_entry_matches_one_device(name, n, nout):
    """ODESBatch_runHostAllDevices: the batch split by index range over every GPU of the box (one plan + one host thread
    per device, no collective) returns what one device returns, bit for bit -- lane layout (MAPK, Goodwin) and the
    warp-per-trajectory layout (huang96, BASELINE configs[4] "at 2/4/8 GPUs") -- and what the reference's CVODE returns.
    On a one-GPU box it degenerates to one range."""
    L = so.lib()
    w = H.workload(name, nout=nout)
    m = w["model"]
    vals = np.ascontiguousarray(H.workload_values(w, n, seed=21))
    vary = np.asarray(w["vary"], dtype=np.int32)
    store = np.arange(m.neq, dtype=np.int32)
    b = so.Batch(m, w["opt"], w["vary"], store=list(store))
    b.set_values(vals)
    b.integrate(n)
    direct = np.transpose(b.results(n), (2, 0, 1)).copy()
    b.close()
    ns = min(n, 512)
    ref = _oracle(w, vals[:ns])
    assert np.array_equal(direct[:ns], ref["out"][:, :, :m.neq], equal_nan=True)
    ndev = L.ODES_deviceCount()
    for want in sorted({1, min(2, ndev), ndev}):
        res = np.full((n, nout + 1, m.neq), -1.0)
        st = np.full(n, -1, dtype=np.int32)
        used = L.ODESBatch_runHostAllDevices(m.om, w["opt"], len(vary), vary.ctypes.data, len(store), store.ctypes.data, n,
                                             vals.ctypes.data, res.ctypes.data, st.ctypes.data, want)
        assert used == want, so.errors()
        assert (st == 0).all() and np.array_equal(res, direct)


@pytest.mark.parametrize("warp", [0, 1])
def test_edge_sizes_through_the_host_entries(monkeypatch, warp):
    """Empty, single and ragged batches through ODESBatch_runHost and ODESBatch_runHostAllDevices, both layouts."""
    monkeypatch.setenv("ODES_WARP", str(warp))
    L = so.lib()
    w = H.workload("mapk", nout=8)
    m = w["model"]
    vals = np.ascontiguousarray(H.workload_values(w, 70, seed=31))
    vary = np.asarray(w["vary"], dtype=np.int32)
    ref = np.transpose(_oracle(w, vals)["out"], (0, 1, 2))            # [set][row][value]
    b = so.Batch(m, w["opt"], w["vary"])
    for n in (0, 1, 33, 70):
        res = np.full((max(n, 1), 9, m.nvalues), -1.0)
        st = np.full(max(n, 1), -1, dtype=np.int32)
        b.run_host(vals[:n], res, st, chunk=0)
        assert np.array_equal(res[:n], ref[:n]) and (st[:n] == 0).all(), n
        res[:] = -1.0
        st[:] = -1
        used = L.ODESBatch_runHostAllDevices(m.om, w["opt"], len(vary), vary.ctypes.data, 0, None, n, vals.ctypes.data, res.ctypes.data, st.ctypes.data, 1)
        assert used == 1 and np.array_equal(res[:n], ref[:n]) and (st[:n] == 0).all(), n
    b.close()


def test_time_dependent_model_on_gpu(tmp_path):
    """Explicit time in a rule and in an event trigger (time as float), exp and ln in the formulas:
    both are evaluated as the host libm evaluates them (glibc_pow.inc), so the trajectories are bit-identical to the
    oracle here too; the event fires at the same output row."""
    from test_frontend_models import TIME_MODEL
    p = tmp_path / "forced.xml"
    p.write_text(TIME_MODEL)
    m = so.Model(str(p))
    vary = [m.index("w"), m.index("k"), m.index("x")]
    rng = np.random.default_rng(8)
    vals = np.stack([rng.uniform(0.3, 2.0, 500), rng.uniform(0.2, 3.0, 500), rng.uniform(0.5, 3, 500)], axis=1)
    opt = so.settings(7.0, 70, 1e-9, 1e-6)
    b = so.Batch(m, opt, vary)
    b.set_values(vals)
    b.integrate(500)
    out, st, fired = b.results(500), b.status(500), b.event_log(500).T
    ref = H.oracle_run(m, opt, vary, vals, 70, compiled_so=H.compile_c_model(m, "forced"), threads=8, backend=H.gpu_tier_backend())
    assert (st == 0).all() and np.array_equal(fired, ref["fired"])
    assert np.array_equal(b.stats(500).T, ref["stats"])
    assert np.array_equal(out, np.transpose(ref["out"], (1, 2, 0)))
    b.close()


@pytest.mark.parametrize("name,nsets,nout", [("mapk", 512, None), ("goodwin", 128, None), ("events", 256, None), ("huang", 64, None)])
def test_adams_moulton_on_gpu(name, nsets, nout):
    """CvodeMethod 1 (Adams-Moulton, orders 1..12) in the warp kernel: bit-identical to the vendored CVODE with CV_ADAMS."""
    w = H.workload(name, nout=nout, method=1)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=2028)
    got = _gpu_run(w, vals)
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF, compiled_so=H.compile_c_model(m, name), threads=8, adams=True)
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["stats"], ref["stats"])
    assert np.array_equal(got["fired"], ref["fired"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)


@pytest.mark.parametrize("name,method,nsets", [("goodwin", 1, 128), ("goodwin", 0, 128), ("events", 1, 256)])
def test_functional_iteration_on_gpu(name, method, nsets):
    """IterMethod 1 (CVnlsFunctional) with BDF or Adams-Moulton in the warp kernel: bit-identical to the vendored CVODE."""
    w = H.workload(name, method=method, iter_method=1)
    m = w["model"]
    vals = H.workload_values(w, nsets, seed=2029)
    got = _gpu_run(w, vals)
    ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF, compiled_so=H.compile_c_model(m, name), threads=8, adams=bool(method), functional=True)
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["stats"], ref["stats"])
    assert np.array_equal(got["fired"], ref["fired"])
    assert np.array_equal(got["out"], np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)


def _refined_settings(w, k):
    """The print grid of a workload with k equal sub-steps per print step, as a designed time series (what the kernel walks
    with StepResolution = k)."""
    tp = H.tpoints(w["opt"], w["nout"])
    grid = np.empty(w["nout"] * k + 1)
    for i in range(w["nout"]):
        for j in range(k):
            grid[i * k + j] = tp[i] if j == 0 else tp[i] + j * ((tp[i + 1] - tp[i]) / k)
    grid[-1] = tp[-1]
    opt = so.settings(w["end"], w["nout"] * k, w["atol"], w["rtol"])
    arr = (C.c_double * (len(grid) - 1))(*grid[1:])
    assert so.lib().CvodeSettings_setTimeSeries(opt, arr, len(grid) - 1) == 1
    return opt, grid


@pytest.mark.parametrize("warp", [0, 1])
def test_step_resolution_refines_event_detection(monkeypatch, warp):
    """cvodeSettings_t.StepResolution = k (reference integratorSettings.h:61-63, "useful for fine-grained event detection";
    the reference never reads it): the solver is driven over k sub-steps per print step and events are examined, assignments
    applied and the solver re-initialised at every sub-step with the reference's output-time semantics; rows are kept at the
    print steps.  Checked against the reference's CVODE run ON THE REFINED GRID: the kept rows are its rows k, 2k, ...
    bit for bit, and the event mask of a kept row is the OR of its masks since the row before."""
    monkeypatch.setenv("ODES_WARP", str(warp))
    L = so.lib()
    k = 4
    w = H.workload("events", nout=25)
    m = w["model"]
    vals = H.workload_values(w, 256, seed=5)
    opt_k = so.settings(w["end"], w["nout"], w["atol"], w["rtol"])
    so.lib().CvodeSettings_setStepResolution(opt_k, k)
    b = so.Batch(m, opt_k, w["vary"])
    assert ("-DODES_WARP=1" in b.source().split("\n")[0]) == bool(warp)
    b.set_values(vals)
    b.integrate(256)
    out, fired, st = b.results(256), b.event_log(256), b.status(256)
    b.close()
    opt_ref, grid = _refined_settings(w, k)
    ref = H.oracle_run(m, opt_ref, w["vary"], vals, w["nout"] * k, backend=H.gpu_tier_backend(), compiled_so=H.compile_c_model(m, "events"), threads=8)
    assert (st == 0).all() and (ref["status"] == 0).all()
    want = np.transpose(ref["out"], (1, 2, 0))[::k]
    assert np.array_equal(out, want, equal_nan=True)
    acc = np.zeros((w["nout"] + 1, 256), dtype=np.uint32)
    for r in range(1, w["nout"] + 1):
        acc[r] = np.bitwise_or.reduce(ref["fired"][:, (r - 1) * k + 1:r * k + 1], axis=1)
    assert np.array_equal(fired, acc)
    # the refinement matters: on the print grid alone some events are seen later (different rows)
    b1 = so.Batch(m, w["opt"], w["vary"])
    b1.set_values(vals)
    b1.integrate(256)
    coarse = b1.results(256)
    b1.close()
    assert not np.array_equal(coarse, out, equal_nan=True)
    # halting modes are refused with StepResolution > 1
    opt_h = so.settings(w["end"], w["nout"], w["atol"], w["rtol"], halt_on_event=1)
    so.lib().CvodeSettings_setStepResolution(opt_h, k)
    with pytest.raises(so.OdesError):
        so.Batch(m, opt_h, w["vary"])
    L.SolverError_clear()


def test_single_instance_without_event_processing():
    """IntegratorInstance_integrateOneStepWithoutEventProcessing (src/integratorInstance.c:1634-1642): the device course of a
    single instance skips the model's event function, so the plain decay S1 -> S2 follows exp(-t); a following
    IntegratorInstance_integrateOneStep switches event handling back on from the current state."""
    L = so.lib()
    m = so.Model(H.model_path("two_events_one_assignment.xml"))
    opt = so.settings(10.0, 100, 1e-12, 1e-9)
    ii = L.IntegratorInstance_create(m.om, opt)
    assert ii
    vi = L.ODEModel_getVariableIndex(m.om, b"S1")
    for k in range(1, 61):
        assert L.IntegratorInstance_integrateOneStepWithoutEventProcessing(ii) == 1
        t = L.IntegratorInstance_getTime(ii)
        assert abs(L.IntegratorInstance_getVariableValue(ii, vi) - np.exp(-t)) < 1e-7, (k, t)
    assert L.IntegratorInstance_getVariableValue(ii, vi) < 0.01           # far below the trigger threshold 0.1: no event was applied
    assert L.IntegratorInstance_integrateOneStep(ii) == 1                  # with events: S1 < 0.1 fires and S1 is set back to 1
    assert abs(L.IntegratorInstance_getVariableValue(ii, vi) - 1.0) < 1e-12
    L.IntegratorInstance_free(ii)


@pytest.mark.parametrize("warp", [0, 1])
def test_event_location_inside_steps_on_gpu(monkeypatch, warp):
    """StepResolution < 0 on the device (see tests/test_emulated_kernel.py::test_emulated_event_location_inside_steps): rows,
    event masks and work counters of 512 sets bit-identical to the vendored CVODE driven step by step through the same
    location procedure."""
    monkeypatch.setenv("ODES_WARP", str(warp))
    nout = 50
    w = H.workload("events", nout=nout)
    m = w["model"]
    vals = H.workload_values(w, 512, seed=9)
    opt = so.settings(w["end"], nout, w["atol"], w["rtol"])
    so.lib().CvodeSettings_setStepResolution(opt, -1)
    b = so.Batch(m, opt, w["vary"])
    head = b.source().split("\n")[0]
    assert "-DODES_ROOTFIND=1" in head and ("-DODES_WARP=1" in head) == bool(warp)
    b.set_values(vals)
    b.integrate(512)
    out, fired, st, stats = b.results(512), b.event_log(512), b.status(512), b.stats(512).T
    b.close()
    ref = H.oracle_run(m, opt, w["vary"], vals, nout, backend=H.REF, compiled_so=H.compile_c_model(m, "events"), threads=8)
    assert (st == 0).all() and (ref["status"] == 0).all()
    assert np.array_equal(out, np.transpose(ref["out"], (1, 2, 0)), equal_nan=True)
    assert np.array_equal(fired.T, ref["fired"]) and ref["fired"].any()
    assert np.array_equal(stats, ref["stats"])
