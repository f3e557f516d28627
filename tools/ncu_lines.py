"""Per-source-line / per-function sample shares from `ncu --page source` exports (development tool):
   ncu -i rep --page source --print-source cuda,sass --csv > a.csv; ncu -i rep --page source --csv > b.csv; python tools/ncu_lines.py a.csv b.csv"""
import bisect, collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
addr2src, file, line = {}, None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path": file = r[1].split('/')[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != '':
        try: line = int(r[0])
        except ValueError: pass
    if len(r) > 3 and r[2].startswith('0x'): addr2src[int(r[2], 16)] = (file, line)
srows = list(csv.reader(open(sys.argv[2])))
h = srows[1]; ix = {k: i for i, k in enumerate(h)}; data = srows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except (ValueError, KeyError): return 0.0
def outline(path):
    st, txt = [], open(path).read().split('\n')
    for i, l in enumerate(txt, 1):
        if l.startswith('template'): continue
        if re.match(r'^(__device__|__global__|static|inline)', l):
            m = re.search(r'([A-Za-z_0-9]+)\s*\(', l)
            if m: st.append((i, m.group(1)))
    return st
import os
base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mpp_cnn_rs_object_detection_b200', 'csrc') + os.sep
out = {n: outline(base + n) for n in ('mpp_sweep2.cuh', 'mpp_device.cuh', 'mpp_proposals.cuh', 'mpp_clip.cuh')}
def fn_of(file, line):
    if file in out and out[file]:
        st = out[file]; i = bisect.bisect_right([s for s, _ in st], line) - 1
        return st[i][1] if i >= 0 else '?'
    return file
tot = sum(f(r, "# Samples") for r in data)
nb = lambda r: f(r, "# Samples") - f(r, "stall_barrier")
tot_nb = sum(nb(r) for r in data)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0]); lines = collections.Counter()
for r in data:
    src = addr2src.get(int(r[0], 16), ('?', 0)); fn = fn_of(*src)
    g = agg[fn]; g[0] += 1; g[1] += f(r, "Instructions Executed"); g[2] += nb(r); g[3] += f(r, "stall_barrier")
    lines[src] += nb(r)
print(f"samples {tot:.0f}, non-barrier {tot_nb:.0f}")
print("%-26s %6s %9s %9s %9s" % ("function", "static", "exec M", "active %", "barrier %"))
for fn, g in sorted(agg.items(), key=lambda x: -x[1][2])[:28]:
    print("%-26s %6d %9.1f %9.2f %9.2f" % (fn, g[0], g[1] / 1e6, 100 * g[2] / tot_nb, 100 * g[3] / max(1, tot - tot_nb)))
print("top lines (non-barrier samples):")
for (fl, ln), v in lines.most_common(30): print(f"  {100 * v / tot_nb:5.2f}%  {fl}:{ln}")
