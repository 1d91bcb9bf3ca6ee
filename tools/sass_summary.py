"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (and of legacy mma.sync HMMA) in the
in-tree libgdf_b200.so: `python tools/sass_summary.py > profiles/rNN_sass_summary.md`."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "generic_diffusion_feature_b200", "libgdf_b200.so")
MNEM = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA.16816", "MUFU.EX2", "SYNCS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in MNEM:
        if re.search(r"\b" + re.escape(k), line):
            counts[cur][k] += 1
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["instr"] += 1
print("# SASS summary of generic_diffusion_feature_b200/libgdf_b200.so (cuobjdump -sass, sm_100a)\n")
print("| kernel | SASS instr | " + " | ".join(MNEM) + " |")
print("|---|---|" + "---|" * len(MNEM))
tot = collections.Counter()
for name, c in sorted(counts.items(), key=lambda kv: -kv[1]["instr"]):
    tot.update(c)
    if any(c[k] for k in MNEM[:8]):
        print("| `%s` | %d | " % (name[:110], c["instr"]) + " | ".join(str(c[k]) for k in MNEM) + " |")
print("| **all %d kernels** | %d | " % (len(counts), tot["instr"]) + " | ".join(str(tot[k]) for k in MNEM) + " |")
no_tc = [n for n, c in counts.items() if not any(c[k] for k in MNEM[:8])]
print("\nKernels without tensor-core / TMA instructions (element-wise, normalisation, resize, gather): %d\n" % len(no_tc))
print(", ".join("`%s`" % n.replace("gdf::", "")[:60] for n in sorted(set(no_tc))))
