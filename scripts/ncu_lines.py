"""per-source-line dynamic instruction counts of one kernel from an ncu report + the cubin's line table
usage: python scripts/ncu_lines.py <report.ncu-rep> <cubin> <mangled kernel name> [top]"""
import csv, io, re, subprocess, sys
from collections import Counter, defaultdict
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ncu = [(int(r[iex]), int(r[isamp]), r[isrc]) for r in rows[hi + 1:] if len(r) > iex and r[iex].isdigit()]
text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(text) if l.startswith(".text." + kern))
cur, lines = None, []
for l in text[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and lines:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        lines.append((cur, m.group(2)))
assert len(ncu) == len(lines), (len(ncu), len(lines))
tot, tots = sum(x[0] for x in ncu), sum(x[1] for x in ncu)
per, perop, samp, ops = Counter(), defaultdict(Counter), Counter(), Counter()
for (ex, s, src), (ln, op) in zip(ncu, lines):
    o = "IMOV" if op.startswith("IMAD.MOV") else op.split(".")[0]
    per[ln] += ex; perop[ln][o] += ex; samp[ln] += s; ops[o] += ex
print(f"warp instructions executed {tot:.4g}; static {len(ncu)}")
print("opcode mix: " + "  ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(16)))
cache = {}
def srcline(f, n):
    import os
    for d in ("pyro2_b200/csrc/", "/usr/local/cuda/include/", "/usr/local/cuda/include/crt/"):
        if os.path.exists(d + f):
            cache.setdefault(d + f, open(d + f).read().split("\n"))
            return cache[d + f][n - 1].strip()[:80]
    return ""
for ln, ex in per.most_common(top):
    print(f"{ln[0][:22]}:{ln[1]:4d} exec {100 * ex / tot:5.2f}% samp {100 * samp[ln] / tots:5.2f}%  "
          + " ".join(f"{k}:{100 * v / tot:.1f}" for k, v in perop[ln].most_common(4)) + "  | " + srcline(*ln))
