"""Per-kernel histogram of the Blackwell-specific SASS opcodes in the shipped library (no GPU needed):
  python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt
UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4 (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA load),
UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc, HMMA/IMMA = legacy mma.sync
(must be absent), SYNCS = mbarrier ops."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "diffusiondepth_b200", "libddengine.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
OPS = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "HMMA", "IMMA", "SYNCS", "ELECT")
cur, table = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        table[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                table[cur][op if o in ("UTCHMMA", "UTCQMMA", "UTMALDG", "LDTM") else o] += 1
print(f"# {os.path.relpath(lib, ROOT)}  ({os.path.getsize(lib)} bytes); cuobjdump -sass, opcode counts per kernel (static instructions)")
tot = collections.Counter()
for fn, c in table.items():
    if not c:
        continue
    name = demangle(fn)
    name = (name[:name.index(">(") + 1] if ">(" in name else re.sub(r"\(.*", "", name)).replace("(int)", "").replace("(bool)", "")
    print(f"{name}\n    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
    tot.update(c)
print("\n# library totals\n    " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items())))
print("# legacy tensor-core opcodes (HMMA / IMMA):", tot.get("HMMA", 0) + tot.get("IMMA", 0))
