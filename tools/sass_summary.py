"""Writes profiles/sass_summary.txt: for every kernel in libais_b200.so the ptxas resource line (registers, spills, shared
memory) from build_ptxas.log and the count of the SASS opcodes that prove which hardware path a kernel uses
(UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = TMEM loads/stores, UTMALDG/UBLKCP = TMA, HMMA = mma.sync, DFMA/DMUL = fp64).

    python tools/sass_summary.py            # needs nvcc's cuobjdump; runs on the CPU-only container
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "anime-illust-image-searcher_b200")
SO = os.path.join(PKG, "libais_b200.so")
LOG = os.path.join(PKG, "build_ptxas.log")
OUT = os.path.join(ROOT, "profiles", "sass_summary.txt")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "DFMA", "DMUL", "DADD",
         "FFMA", "REDUX", "LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "BAR"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:   # noqa: BLE001
        return {n: n for n in names}


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    res = {}
    cur = None
    for ln in open(LOG):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
        if m:
            cur = m.group(1)
            res[cur] = {}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m:
            res[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", ln)
        if m:
            res[cur]["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", ln)
            res[cur]["smem"] = int(s.group(1)) if s else 0
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    ops = collections.defaultdict(collections.Counter)
    fn = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and fn:
            ops[fn][m.group(1)] += 1
    names = sorted(set(res) | set(ops))
    dm = demangle(names)
    lines = ["# libais_b200.so, sm_100a: ptxas resources + SASS opcode counts per kernel (tools/sass_summary.py)",
             "# columns: regs  smem(B)  stack/spill-st/spill-ld (B)  total-instr  watched opcodes",
             ""]
    for n in sorted(names, key=lambda x: short(dm[x])):
        r = res.get(n, {})
        c = ops.get(n, collections.Counter())
        total = sum(c.values())
        watch = " ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])
        lines.append("%-58s regs=%-3s smem=%-6s stack/spill=%s/%s/%s instr=%-5d %s" % (
            short(dm[n])[:58], r.get("regs", "?"), r.get("smem", "?"), r.get("stack", "?"), r.get("spill_st", "?"),
            r.get("spill_ld", "?"), total, watch))
    spills = [short(dm[n]) for n in names if res.get(n, {}).get("spill_st", 0) or res.get(n, {}).get("spill_ld", 0)]
    lines += ["", "kernels with register spills: %s" % (", ".join(spills) if spills else "none")]
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    sys.exit(main())
