#!/usr/bin/env python
"""Counts the SASS mnemonics that prove which hardware paths the built library uses (B200_PROFILING.md):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk), DMMA (mma.sync f64), SYNCS (mbarrier), per kernel.    python profiles/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ann_force_b200", "libannforce_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
counts, order, name = collections.defaultdict(collections.Counter), [], None
pats = {"UTCHMMA.2CTA": r"\bUTCHMMA\.2CTA", "UTCHMMA": r"\bUTCHMMA\b(?!\.2CTA)", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
        "UBLKCP": r"\bUBLKCP", "DMMA": r"\bDMMA", "SYNCS": r"\bSYNCS", "UTCBAR": r"\bUTCBAR", "total": r"^\s+/\*[0-9a-f]{4,}\*/"}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); order.append(name); continue
    if name:
        for k, pat in pats.items():
            if re.search(pat, line): counts[name][k] += 1
names = demangle(order)
print("# SASS mnemonic counts per kernel of ann_force_b200/libannforce_b200.so (cuobjdump -sass; sm_100a)")
print("# %-86s %s" % ("kernel", " ".join("%12s" % k for k in pats)))
tot = collections.Counter()
for raw, nice in zip(order, names):
    c = counts[raw]
    if not any(c[k] for k in pats if k != "total" and k != "SYNCS"): continue
    nice = re.sub(r"\(.*", "", nice).replace("annf::", "").replace("void ", "")
    print("  %-86s %s" % (nice[:86], " ".join("%12d" % c[k] for k in pats)))
    tot.update(c)
print("  %-86s %s" % ("ALL KERNELS LISTED", " ".join("%12d" % tot[k] for k in pats)))
