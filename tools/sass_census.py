"""Per-kernel census of the Blackwell-specific SASS in libvls_b200.so (cuobjdump -sass): tcgen05 MMAs (UTCHMMA), tensor-memory
loads / stores (LDTM / STTM), TMA loads / stores (UTMALDG / UTMASTG), tcgen05 commits (UTCBAR), warp-level mma.sync / movmatrix of the decoder's token-side cluster kernel (HMMA / MOVM), mbarrier ops (SYNCS), SFU
exponentials (MUFU.EX2), packed FP32 (FFMA2 / FADD2) and the ELECT + BRA.U.ANY loops the compiler emits around uniform-datapath
instructions in divergent code (must be 0).  usage: python tools/sass_census.py > profiles/sass_census.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "video-llava-seg_b200", "libvls_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTCHMMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
                               ("UTMASTG", r"\bUTMASTG\b"), ("UTCBAR", r"\bUTCBAR\b"), ("HMMA", r"\bHMMA\b"), ("MOVM", r"\bMOVM\b"), ("SYNCS", r"\bSYNCS\b"), ("MUFU.EX2", r"MUFU\.EX2"),
                               ("FFMA2", r"\bFFMA2\b"), ("FADD2", r"\bFADD2\b"), ("BRA.U.ANY", r"BRA\.U\.ANY"), ("instr", r"^\s+/\*[0-9a-f]{4,6}\*/")])
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for k, p in PAT.items():
        if re.search(p, line):
            counts[name][k] += 1
print(f"# SASS census of {os.path.relpath(so, ROOT)} (sm_100a), {len(counts)} kernels; columns: " + " ".join(PAT))
rows = []
for n, c in counts.items():
    d = re.sub(r"\(anonymous namespace\)::|vls::", "", demangle(n))
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    rows.append((d, c))
for d, c in sorted(rows, key=lambda r: (-r[1]["UTCHMMA"], r[0])):
    print(f"{d[:64]:64s} " + " ".join(f"{c[k]:6d}" for k in PAT))
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
print(f"{'TOTAL':64s} " + " ".join(f"{tot[k]:6d}" for k in PAT))
