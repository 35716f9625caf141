"""Join an ncu SASS-page CSV with nvdisasm line info and aggregate by source line / function.
usage: ncu_by_line.py sass.csv dis.txt kernel_mangled_substring"""
import csv, re, sys, collections
sass_csv, dis, kern = sys.argv[1:4]
# parse disassembly of the kernel section
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and kern in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("\t.section\t.text.")), len(lines))
cur = ("?", 0); func = "entry"; info = {}
for l in lines[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'^(\S+):\s*$', l)
    if m and not m.group(1).startswith(".L_x"):
        func = m.group(1); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m:
        info[int(m.group(1), 16)] = (func, cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
by_line = collections.defaultdict(lambda: [0, 0, 0, 0, 0]); by_func = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]
for r in rows[2:]:
    off = int(r[0], 16) - base
    f, (file, line), txt = info.get(off, ("?", ("?", 0), ""))
    v = [1, int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]]), int(r[ix["stall_no_inst"]]) if "stall_no_inst" in ix else 0]
    for acc in (by_line[(file, line)], by_func[f], tot):
        for k in range(5): acc[k] += v[k]
def show(d, n, title):
    print(title)
    print(f"{'key':70s} {'static':>7s} {'warp_inst%':>10s} {'thr/warp':>8s} {'samples%':>8s} {'noinst%':>8s}")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][3])[:n]:
        print(f"{str(k)[:70]:70s} {v[0]:7d} {100*v[1]/tot[1]:10.2f} {v[2]/max(1,v[1]):8.1f} {100*v[3]/tot[3]:8.2f} {100*v[4]/max(1,v[3]):8.1f}")
print("total static", tot[0], "warp inst", tot[1], "thread inst", tot[2], "samples", tot[3])
show(by_func, 40, "== by function ==")
show(by_line, 60, "== by source line ==")
