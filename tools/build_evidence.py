#!/usr/bin/env python
"""Static evidence of the built library (no GPU needed): `python tools/build_evidence.py r02` writes
profiles/<tag>_ptxas_summary.txt (registers / spills of every kernel, from the -Xptxas -v logs of the last build) and
profiles/<tag>_sass_opcodes.txt (opcode histogram of every kernel in libgen3c_b200.so from `cuobjdump -sass`, with the
Blackwell-specific mnemonics — UTCHMMA[.2CTA], LDTM / STTM, UTMALDG[.MULTICAST], UTMASTG, UTMAREDG, UTCBAR — listed
first)."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCCP", "SYNCS", "REDG", "RED", "MUFU",
       "FFMA2", "FADD2", "FMUL2")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def ptxas_summary():
    lines = ["# nvcc -Xptxas -v for sm_100a (gen3c_b200/csrc/Makefile): registers / spills of every kernel in libgen3c_b200.so"]
    for log in sorted(glob.glob(os.path.join(ROOT, "gen3c_b200", "lib", "obj", "*.ptxas.log"))):
        t = open(log).read()
        recs = re.findall(r"Compiling entry function '(\S+)'.*?\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                          r"(\d+) bytes spill loads\n.*?Used (\d+) registers", t)
        if not recs:
            continue
        dm = demangle([r[0] for r in recs])
        lines.append("## " + os.path.basename(log).replace(".ptxas.log", ".cu"))
        for name, stack, st, ld, regs in recs:
            short = re.sub(r"\(.*", "", dm.get(name, name))
            lines.append(f" {int(regs):3d} regs  spill st/ld {int(st):3d}/{int(ld):3d} B  stack {int(stack):3d} B  {short}")
    return "\n".join(lines) + "\n"


def sass_histogram():
    so = os.path.join(ROOT, "gen3c_b200", "lib", "libgen3c_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, per = None, collections.OrderedDict()
    for ln in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", ln)
        if m and cur:
            op = m.group(1)
            base = op.split(".")[0]
            key = op if base in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "LDTM", "STTM", "MUFU", "RED", "REDG") else base
            per[cur][key] += 1
    dm = demangle(list(per))
    out = [f"# cuobjdump -sass gen3c_b200/lib/libgen3c_b200.so : opcode counts per kernel (static instruction counts)",
           "# Blackwell-native mnemonics first: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st,",
           "# UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce (.MULTICAST across the cluster), UTCBAR = tcgen05.commit"]
    tot = collections.Counter()
    for fn, c in per.items():
        if not c:
            continue
        name = re.sub(r"\(.*", "", dm.get(fn, fn))
        keyops = [(k, v) for k, v in sorted(c.items()) if k.split(".")[0] in KEY]
        rest = [(k, v) for k, v in c.most_common() if k.split(".")[0] not in KEY][:12]
        out.append(f"## {name}  ({sum(c.values())} instructions)")
        out.append("   " + "  ".join(f"{k}:{v}" for k, v in keyops))
        out.append("   " + "  ".join(f"{k}:{v}" for k, v in rest))
        tot.update(c)
    out.insert(3, "## whole library: " + "  ".join(f"{k}:{v}" for k, v in sorted(tot.items()) if k.split('.')[0] in KEY[:9]))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_ptxas_summary.txt"), "w").write(ptxas_summary())
    open(os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.txt"), "w").write(sass_histogram())
    print("wrote profiles/%s_ptxas_summary.txt and profiles/%s_sass_opcodes.txt" % (tag, tag))
