"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by (file, line): instructions executed, samples,
average active threads.  Usage: python scripts/ncu_lines.py export.csv [top_n]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
cur_file = None; hdr = None; cur_line = None
tot_i = tot_s = tot_t = 0
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None: continue
    if row[0] not in ("", "-"):
        cur_line = (cur_file, int(row[0]), row[1].strip()[:90])
    if row[2] in ("-", ""): continue   # source-only row
    try:
        ins = int(row[hdr.index("Instructions Executed")]); thr = int(row[hdr.index("Thread Instructions Executed")]); smp = int(row[hdr.index("# Samples")])
    except ValueError: continue
    a = agg.setdefault(cur_line, [0, 0, 0, 0]); a[0] += ins; a[1] += thr; a[2] += smp; a[3] += 1
    tot_i += ins; tot_t += thr; tot_s += smp
print(f"total warp-instr {tot_i}  thread-instr {tot_t}  avg lanes {tot_t/max(tot_i,1):.2f}  samples {tot_s}")
byfile = collections.Counter()
for (f, l, s), a in agg.items(): byfile[f] += a[0]
print({k: f"{100*v/tot_i:.1f}%" for k, v in byfile.items()})
for (f, l, s), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{100*a[0]/tot_i:5.1f}%i {100*a[2]/max(tot_s,1):5.1f}%s lanes {a[1]/max(a[0],1):5.1f} sass {a[3]:4d}  {f}:{l}  {s}")
