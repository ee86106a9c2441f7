"""SASS opcode summary of libmgf_sm100a.so per kernel family (all template instances summed): the mnemonics that prove the Blackwell
paths -- UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (cp.async.bulk), LDTM (tcgen05.ld), HMMA (mma.sync).
    python scripts/sass_summary.py > profiles/rNN_sass_opcodes.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "morphganformer_b200", "lib", "libmgf_sm100a.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"Function : (\S+)", sass)))
dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()))
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "HMMA", "LDSM", "SHFL", "LDG", "STG", "LDS", "STS", "ATOMS", "RED"]
cnt, tot, inst = collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        d = dem[m.group(1)]
        d = re.sub(r"\(anonymous namespace\)::", "", d).replace("void ", "")
        cur = re.sub(r"[<(].*", "", d).replace("mgf::", "")
        inst[cur] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        tot[cur] += 1
        if m.group(1) in OPS:
            cnt[cur][m.group(1)] += 1
print("SASS of `morphganformer_b200/lib/libmgf_sm100a.so` (sm_100a), opcode counts per kernel family, all template instances summed.\n")
print("| kernel | instances | SASS instr | " + " | ".join(OPS) + " |")
print("|---|---:|---:|" + "---:|" * len(OPS))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("| `%s` | %d | %d | " % (k, inst[k], tot[k]) + " | ".join(str(cnt[k][o]) if cnt[k][o] else "" for o in OPS) + " |")
allc = collections.Counter()
for k in cnt:
    allc.update(cnt[k])
print("\nTotals: " + ", ".join("%s %d" % (o, allc[o]) for o in OPS if allc[o]))
