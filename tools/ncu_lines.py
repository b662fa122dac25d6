"""join an ncu `--page source --csv` export (SASS rows of one kernel) with the line info of the shipped cubin:
   python tools/ncu_lines.py <source.csv> <nvdisasm -g dump> <mangled kernel name> [first data row] [rows]
prints the per-source-line share of executed instructions, of stall samples, and the thread efficiency"""
import collections, csv, re, sys
srcdir = "lammps-sph-multiphase_b200/csrc/"
rows = list(csv.reader(open(sys.argv[1])))
dis = open(sys.argv[2]).read().split("\n")
name = sys.argv[3]
# the csv may hold several kernels: blocks start with a "Kernel Name" row followed by a header row
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
b0 = blocks[which]; b1 = blocks[which + 1] if which + 1 < len(blocks) else len(rows)
H = rows[b0 + 1]; data = rows[b0 + 2:b1]
start = [i for i, l in enumerate(dis) if l.startswith(".text." + name + ":")][0]
cur, ins = None, []
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2), cur))
assert len(ins) == len(data), (len(ins), len(data))
ii, it, isamp = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("# Samples")
stalls = [c for c in H if c.startswith("stall_") and "Not Issued" not in c]
by = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for r, (txt, cur) in zip(data, ins):
    v = by[cur]; v[0] += int(r[ii]); v[1] += int(r[it]); v[2] += int(r[isamp])
    for c in stalls:
        v[3][c] += int(r[H.index(c)])
tot = sum(v[0] for v in by.values()); tots = sum(v[2] for v in by.values())
print(rows[b0][1], "instructions", tot, "samples", tots)
cache = {}
for k, v in sorted(by.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[5]) if len(sys.argv) > 5 else 30]:
    f, ln = k if k else ("?", 0)
    try:
        src = cache.setdefault(f, open(srcdir + f).read().split("\n"))
        text = src[ln - 1].strip()[:80]
    except Exception:
        text = ""
    top = ", ".join("%s %.0f%%" % (c[6:], 100.0 * n / max(v[2], 1)) for c, n in v[3].most_common(3))
    print("%5.1f%% inst %5.1f%% samples eff %4.1f  %s:%d  %-80s | %s" % (100 * v[0] / tot, 100 * v[2] / tots, v[1] / max(v[0], 1), f, ln, text, top))
