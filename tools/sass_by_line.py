"""Join an ncu SASS source-page CSV with nvdisasm line info -> executed warp-instructions per CUDA source line.
usage: sass_by_line.py <ncu_source.csv> <nvdisasm -g -c output> <kernel substring> <n_units>"""
import csv, re, sys, collections
src_csv, sass_txt, kern, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
# parse nvdisasm: track current line per instruction offset within the kernel's .text section
line_of = {}
cur_line = None; in_k = False; off_re = re.compile(r"/\*([0-9a-f]{4,})\*/")
for l in open(sass_txt, errors="ignore"):
    if l.startswith("//--------------------- .text."):
        in_k = kern in l
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = off_re.search(l)
    if m and cur_line:
        line_of[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
ci = {h: i for i, h in enumerate(rows[hi])}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[hi + 1:]:
    try:
        addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        n = int(r[ci["Instructions Executed"]] or 0); thr = int(r[ci["Thread Instructions Executed"]] or 0); smp = int(r[ci["# Samples"]] or 0)
    except ValueError:
        continue
    if base is None:
        base = addr
    ln = line_of.get(addr - base, ("?", 0))
    a = agg[ln]; a[0] += n; a[1] += thr; a[2] += smp
    tot += n
print(f"total warp-inst {tot}  per unit {tot/units:.1f}")
srcs = {}
for (f, ln), (n, thr, smp) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if n / units < 0.5:
        continue
    if f not in srcs:
        try: srcs[f] = open("/root/repo/npore_b200/csrc/" + f).read().split("\n")
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
    print(f"{n/units:7.1f} inst/unit  act {thr/max(n,1):5.1f}  smp {smp:6d} | {f}:{ln}: {text}")
