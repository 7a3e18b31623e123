"""Instruction-cache view of an ncu source-page CSV: how many SASS instructions of a kernel are hot (executed at least
`thr` times), per source function / task, and how many 128-byte lines they occupy.
usage: ncu_hot_code.py <src.csv> <k.sass from nvdisasm -g -c> <kernel mangled substring> [thr]"""
import re, csv, collections, sys
src_csv, sass, kname = sys.argv[1:4]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 1e6
insec = False; line = None; amap = {}
for l in open(sass):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        insec = kname in m.group(1); line = None; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: amap[int(m.group(1), 16)] = (line, m.group(2))
rows = list(csv.reader(open(src_csv))); h = rows[1]; idx = {n: i for i, n in enumerate(h)}
base = None; agg = collections.defaultdict(lambda: [0, 0, 0]); lines = collections.defaultdict(int); data = []
for r in rows[2:]:
    if len(r) < len(h): continue
    a = int(r[idx["Address"]], 16)
    if base is None: base = a
    ln, txt = amap.get(a - base, (None, ""))
    n = int(r[idx["Instructions Executed"]] or 0)
    key = (ln[0], ln[1] // 10 * 10) if ln else ("?", 0)
    agg[key][0] += 1
    if n >= thr: agg[key][1] += 1
    agg[key][2] += int(r[idx["# Samples"]] or 0)
    lines[(a - base) // 128] = max(lines[(a - base) // 128], n)
    data.append((a - base, n))
hot = sum(v[1] for v in agg.values())
hot_lines = [k for k, v in lines.items() if v >= thr]
cold_in_hot = sum(1 for off, n in data if lines[off // 128] >= thr and n < thr / 10)
print("hot instructions (>= %.0e executions): %d = %.1f KB; 128-byte lines holding them: %d = %.1f KB; cold instructions inside those lines: %d"
      % (thr, hot, hot * 16 / 1024, len(hot_lines), len(hot_lines) * 128 / 1024, cold_in_hot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    if v[1]:
        print("%-26s line %4d..  static %4d  hot %4d  samples %6d" % (k[0], k[1], v[0], v[1], v[2]))
