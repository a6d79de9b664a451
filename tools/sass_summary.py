"""Counts the SASS mnemonics that matter for the evidence trail in the cubins NVRTC produced for sm_100a:
tensor-memory moves (LDTM / STTM, UTCATOMSWS = tcgen05.alloc/dealloc), FP64 arithmetic, shared-memory traffic, total size.
usage: sass_summary.py name=cubin [name=cubin ...]   (cubins: ODES_B200_CACHE=<dir> python tools/dump_source.py ...)"""
import collections, re, subprocess, sys
for spec in sys.argv[1:]:
    name, path = spec.split("=", 1)
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    ops = collections.Counter()
    fn = None; per = collections.defaultdict(collections.Counter)
    for l in sass.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m: fn = m.group(1); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m: ops[m.group(1)] += 1; per[fn][m.group(1)] += 1
    tot = sum(ops.values())
    print(f"== {name}: {path.split('/')[-1]}  ({tot} SASS instructions in {len(per)} functions; arch sm_100a)")
    for k in ("LDTM", "STTM", "UTCATOMSWS", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "IMAD", "FSEL", "ISETP", "BRA", "BSSY", "BSYNC", "WARPSYNC", "SHFL", "REDUX", "BAR", "CALL"):
        if ops[k]: print(f"  {k:12s} {ops[k]:6d}")
    for f, c in per.items(): print(f"  function {f}: {sum(c.values())} instructions")
