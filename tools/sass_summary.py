"""Static evidence from the built library (no GPU needed): per kernel family, how many SASS instructions of the Blackwell
tensor-core / TMA / tensor-memory / PDL kinds it contains, and ptxas' register / spill / shared-memory figures.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (f16 kinds), UTCQMMA / UTCIMMA other kinds, UTMALDG / UTMASTG = TMA tensor
load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, PREEXIT = griddepcontrol.launch_dependents,
ACQBULK = griddepcontrol.wait.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "transception_b200", "libtransception_sm100.so")
KINDS = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "PREEXIT", "ACQBULK", "HMMA", "FFMA", "MUFU.EX2")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_n"] += 1
            for k in KINDS:
                if op == k or op.startswith(k + "."):
                    cur[k] += 1
    dm = demangle(list(per))
    fam = collections.OrderedDict()
    for name, c in per.items():
        d = dm.get(name, name)
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        base = re.sub(r"^void\s+", "", d).split("(")[0]
        f = fam.setdefault(base, collections.Counter())
        f.update(c)
        f["_variants"] += 0
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and fn:
            usage[fn] = tuple(int(x) for x in m.groups())
    print("# %s" % os.path.relpath(LIB, ROOT))
    print("# kernels: %d   (cuobjdump -sass / -res-usage; counts are static instruction counts, not executed counts)" % len(per))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# totals: " + "  ".join("%s=%d" % (k, tot[k]) for k in KINDS if tot[k]))
    print("%-64s %6s %5s %7s %5s  %s" % ("kernel", "instr", "regs", "smem", "lmem", "tensor / TMA / TMEM / PDL instructions"))
    for name, c in per.items():
        d = re.sub(r"\(anonymous namespace\)::", "", dm.get(name, name))
        base = re.sub(r"^void\s+", "", d).split("(")[0]
        u = usage.get(name, (0, 0, 0))
        marks = "  ".join("%s=%d" % (k, c[k]) for k in KINDS if c[k] and k not in ("FFMA",))
        print("%-64s %6d %5d %7d %5d  %s" % (base[:64], c["_n"], u[0], u[1], u[2], marks))
    no_wait = [re.sub(r"\(anonymous namespace\)::", "", dm[n]).split("(")[0] for n, c in per.items() if c["PREEXIT"] and not c["ACQBULK"]]
    print("# kernels that trigger dependents but never wait: %s" % (", ".join(no_wait) if no_wait else "none"))


if __name__ == "__main__":
    sys.exit(main())
