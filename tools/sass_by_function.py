"""Developer tool: join an `ncu --page source --csv` per-instruction table of one kernel with `nvdisasm -g` line info of the same build and
print executed instructions / samples per source function (inlined code is attributed to the function whose body the line lies in).
   cuobjdump -xelf all tray_rust_b200/lib/libtrb.so; nvdisasm -g -c trb_api.sm_100a.cubin > dis.txt
   python tools/sass_by_function.py profile.csv dis.txt <mangled kernel name prefix> [--lines]"""
import bisect, csv, re, sys, os
prof_path, dis_path, kernel = sys.argv[1:4]
by_line = "--lines" in sys.argv
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(prof_path)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
prof = [r for r in rows[2:] if len(r) > col['Source']]
lines = open(dis_path).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kernel))
cur = None; dis = []
for l in lines[start + 1:]:
    if l.startswith('//-----'): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', l): dis.append(cur)
assert len(dis) == len(prof), (len(dis), len(prof))
def func_table(path):
    tab = []
    for n, l in enumerate(open(path), 1):
        m = re.match(r'^(?:template.*>\s*)?(?:__device__|__global__|TRB_HD|TRB_DM|__host__ __device__).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', l)
        if m and not l.startswith(' '): tab.append((n, m.group(1)))
    return tab
tabs = {f: func_table(os.path.join(REPO, 'tray_rust_b200', 'csrc', f)) for f in ('trb_kernels.cuh', 'trb_detmath.cuh', 'trb_anim.h')}
src = {f: open(os.path.join(REPO, 'tray_rust_b200', 'csrc', f)).read().split('\n') for f in tabs}
def key(loc):
    if loc is None: return '?'
    f, l = loc
    if by_line: return '%s:%d %s' % (f, l, src[f][l - 1].strip()[:90] if f in src else '')
    t = tabs.get(f)
    if not t: return f
    i = bisect.bisect_right([x[0] for x in t], l) - 1
    return f.split('.')[0][4:] + ':' + (t[i][1] if i >= 0 else '?')
agg = {}; E = S = 0
for r, loc in zip(prof, dis):
    a = agg.setdefault(key(loc), [0, 0, 0, 0, 0, 0])
    e = int(r[col['Instructions Executed']]); s = int(r[col['# Samples']])
    a[0] += e; a[1] += int(r[col['Thread Instructions Executed']]); a[2] += s
    a[3] += int(r[col['stall_long_sb']]); a[4] += int(r[col['stall_no_inst']]); a[5] += int(r[col['stall_wait']])
    E += e; S += s
print('%-60s %7s %6s %8s %7s %7s %7s' % ('where', 'instr%', 'lanes', 'samples%', 'longsb%', 'noinst%', 'wait%'))
for k, v in sorted(agg.items(), key=lambda x: -x[1][2])[:40]:
    print('%-60s %7.1f %6.1f %8.1f %7.1f %7.1f %7.1f' % (k[:60], 100 * v[0] / E, v[1] / max(v[0], 1), 100 * v[2] / S, 100 * v[3] / S, 100 * v[4] / S, 100 * v[5] / S))
