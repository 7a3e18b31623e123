"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line using nvdisasm -g line info.
usage: ncu_lines.py <src.csv> <k.sass> <kernel mangled substring> [top]"""
import re, collections, csv, sys
src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
insec = False; line = None; amap = {}
for l in open(sass):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        insec = kname in m.group(1); line = None; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        line = (m.group(1).split('/')[-1], int(m.group(2)), tuple((a.split('/')[-1], int(b)) for a, b in inl)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: amap[int(m.group(1), 16)] = (line, m.group(2))
rows = list(csv.reader(open(src_csv)))
h = rows[1]; idx = {n: i for i, n in enumerate(h)}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) < len(h): continue
    a = int(r[idx["Address"]], 16)
    if base is None: base = a
    off = a - base
    ln, txt = amap.get(off, (None, ""))
    key = None
    if ln:
        # attribute to the outermost frame in pcb_trace.cuh when inlined, else the line itself
        key = (ln[0], ln[1])
        for f, n in ln[2]:
            key = (f, n)
    smp = int(r[idx["# Samples"]] or 0); n = int(r[idx["Instructions Executed"]] or 0); t = int(r[idx["Thread Instructions Executed"]] or 0)
    g = agg[key]; g[0] += smp; g[1] += n; g[2] += t
    for c in stall_cols:
        v = int(r[idx[c]] or 0)
        if v: g[3][c] += v
    tot_s += smp; tot_i += n
print("samples", tot_s, "warp-inst", tot_i)
for key, g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(g[0], 1)) for k, v in g[3].most_common(3))
    print("%-28s %6.2f%% smp %6.2f%% inst  thr/inst %5.1f  | %s" % (key, 100 * g[0] / tot_s, 100 * g[1] / tot_i, g[2] / max(g[1], 1), st))
