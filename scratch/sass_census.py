"""SASS census of every object of libcfm_b200 (run on the build box: cuobjdump only): which tensor-core / TMA / async
mnemonics each translation unit contains.  Usage: python scratch/sass_census.py > profiles/r02_sass_census.md"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "SYNCS",
        "ELECT", "SHFL", "MUFU.EX2", "LDS", "STS", "LD.E", "ST.E", "STL", "LDL"]
print("# SASS census of libcfm_b200 (sm_100a), round 2\n")
print("`cuobjdump -sass` of every object under `ceo-recommender_b200/lib/`; counts of instructions whose mnemonic starts")
print("with the column name.  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,")
print("HMMA/LDSM = legacy mma.sync / ldmatrix, LDGSTS = cp.async, LDS/STS = shared-memory accesses, LD.E/ST.E = GENERIC")
print("accesses (should be ~0: a generic access to shared memory is what round 2 found and removed), STL/LDL = register spills.\n")
print("| object | kernel | " + " | ".join(WANT) + " |")
print("|---|---|" + "---:|" * len(WANT))
for obj in sorted(glob.glob(os.path.join(ROOT, "ceo-recommender_b200", "lib", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn).replace("void ", "").replace("cfm::", "")
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and fn:
            op = m.group(1)
            for w in WANT:
                if op.startswith(w):
                    counts[fn][w] += 1
    for fn, c in counts.items():
        if fn.startswith("cub::") or "at::" in fn:
            fn = fn[:60] + "..."
        if sum(c.values()) == 0 and not fn.startswith("tower"):
            continue
        print(f"| {os.path.basename(obj)} | `{fn[:90]}` | " + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
