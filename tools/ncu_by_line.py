"""Dev tool: join an ncu source-page CSV (SASS order) with nvdisasm -g line info -> per-source-line totals.
usage: ncu_by_line.py <rep.ncu-rep> <lib.so> <kernel-symbol-substring> [hops]"""
import csv, re, subprocess, sys, os, tempfile, glob
from collections import defaultdict
rep, lib, sym = sys.argv[1:4]
hops = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
for cubin in sorted(glob.glob(tmp + "/*.cubin")):  # one cubin per translation unit
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    if any(".section" in l and ".text." in l and sym in l for l in dis):
        break
start = [i for i, l in enumerate(dis) if ".section" in l and ".text." in l and sym in l][0]
ends = [i for i, l in enumerate(dis) if i > start and ".section" in l]
end = ends[0] if ends else len(dis)
cur, ins = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
h = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[h]; ix = {x: i for i, x in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
assert len(data) == len(ins), (len(data), len(ins))
agg = defaultdict(lambda: [0, 0])
for src, r in zip(ins, data):
    agg[src][0] += int(r[ix["# Samples"]] or 0)
    agg[src][1] += int(r[ix["Instructions Executed"]] or 0)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
srcs = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:50]:
    f, ln = k if k else ("?", 0)
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "turdb_b200", "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else None
    code = srcs[f][ln - 1].strip()[:64] if srcs[f] else ""
    print("%5.1f%% samp %5.1f%% inst %7.1f inst/hop  %s:%d  %s" % (100 * v[0] / ts, 100 * v[1] / ti, v[1] / hops, f, ln, code))
print("total inst/hop", ti / hops)
