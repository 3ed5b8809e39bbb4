"""SASS evidence for profiles/: per kernel of libcbo_b200.so the counts of the mnemonics that matter on sm_100a (DMMA.8x8x4,
UBLKCP = TMA bulk copy, SYNCS = mbarrier ops, USETMAXREG, LDS, LDG/STG, local-memory spills) and, for the two DMMA kernels
of the prior, the steady-state consumer loop (the instructions between two mbarrier waits with the most DMMAs).
    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt        (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = os.path.join(ROOT, "cbo_with_oop_b200", "libcbo_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for ln in sass.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur:
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m: funcs[cur].append(m.group(1).strip())
def demangle(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:150]
    except Exception: return n
keys = [("DMMA.8x8x4", r"^(@!?U?P\d+\s+)?DMMA\.8x8x4"), ("UBLKCP (TMA bulk)", r"UBLKCP"), ("UTMALDG (TMA tensor map)", r"UTMALDG"), ("SYNCS (mbarrier)", r"SYNCS"),
        ("USETMAXREG", r"USETMAXREG"), ("LDS", r"^(@!?U?P\d+\s+)?LDS"), ("LDG", r"^(@!?U?P\d+\s+)?LDG"), ("STG", r"^(@!?U?P\d+\s+)?STG"),
        ("LDGSTS (cp.async)", r"LDGSTS"), ("LDL/STL (spills)", r"^(@!?U?P\d+\s+)?(LDL|STL)"), ("UTC*MMA / LDTM (tcgen05)", r"UTC.*MMA|LDTM")]
print("SASS of cbo_with_oop_b200/libcbo_b200.so (sm_100a), mnemonic counts per kernel\n")
for name, ins in funcs.items():
    d = demangle(name)
    if "cbo::" not in d: continue
    cnt = {k: sum(1 for t in ins if re.search(p, t)) for k, p in keys}
    print(d); print("   ", len(ins), "instructions;", ", ".join(f"{k}: {v}" for k, v in cnt.items() if v)); print()
for pat in ("prior_eval_kernel", "prior_pair_kernel"):
    for name, ins in funcs.items():
        d = demangle(name)
        if pat not in d or ("prior_eval_kernel" in pat and "false" not in d and ", 0>" not in d and "0)" not in d): continue
        idx = [i for i, t in enumerate(ins) if "SYNCS.PHASECHK" in t]
        best = max(zip(idx, idx[1:] + [len(ins)]), key=lambda ab: sum("DMMA" in t for t in ins[ab[0]:ab[1]]), default=None)
        if not best: continue
        a, b = best
        last = max(i for i in range(a, b) if "DMMA" in ins[i])
        print(f"---- steady-state consumer loop of {d}\n     ({sum('DMMA' in t for t in ins[a:last+1])} DMMA.8x8x4, "
              f"{sum(t.startswith('LDS') for t in ins[a:last+1])} LDS between two mbarrier waits)")
        for t in ins[max(0, a - 2):last + 8]: print("      ", t)
        print()
        break
