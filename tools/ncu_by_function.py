"""Attributes executed instructions / stall samples of an ncu capture to source functions and lines of the
generated translation unit.  usage: ncu_by_function.py prof.ncu-rep model_tu.cu cubin [nlines]
(the cubin must be the one the capture ran: build it here with ODES_B200_CACHE=<dir> python tools/dump_source.py)"""
import collections, csv, io, re, subprocess, sys
rep, tu, cubin = sys.argv[1:4]
nlines = int(sys.argv[4]) if len(sys.argv) > 4 else 25
sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
off2line = {}; cur = None; func = None; kernel_off = {}
for l in sass.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: func = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m and func == "odes_bdf_kernel": off2line[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
base = int(rows[2][0], 16)
src = open(tu).read().split("\n")
funcs = [(i, l.strip()[:70]) for i, l in enumerate(src, 1) if re.match(r"\s*(template\s*<[^>]*>\s*)?(DEV|ODES_NOINLINE) ", l) or "__global__" in l]
funcs.append((len(src) + 1, "END"))
def fn(ln):
    if ln is None: return "(other functions: f_eval, root_pow, division slow paths)"
    for k in range(len(funcs) - 1):
        if funcs[k][0] <= ln < funcs[k + 1][0]: return funcs[k][1]
    return "pre"
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); byline = collections.defaultdict(lambda: [0, 0, 0, 0])
for r in rows[2:]:
    try: n = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except Exception: continue
    ln = off2line.get(int(r[0], 16) - base)
    for d, k in ((agg, fn(ln)), (byline, ln)):
        a = d[k]; a[0] += n; a[1] += t; a[2] += s; a[3] += 1
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("exec%  samples%  avg_threads  static  function")
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:30]:
    print(f"{100 * a[0] / tot:5.1f}% {100 * a[2] / ts:7.1f}% {a[1] / max(a[0], 1):10.1f} {a[3]:7d}  {f}")
print("\ntop lines by stall samples")
for ln, a in sorted(byline.items(), key=lambda kv: -kv[1][2])[:nlines]:
    print(f"{100 * a[2] / ts:5.2f}% exec {100 * a[0] / tot:5.2f}% thr {a[1] / max(a[0], 1):5.1f} static {a[3]:4d} L{ln}: {src[ln - 1].strip()[:100] if ln else '?'}")
