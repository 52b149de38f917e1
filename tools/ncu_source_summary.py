"""Summarise `ncu --page source --csv` of one kernel: stall-reason totals, and the hottest SASS instructions by
sampled stalls and by executed warp instructions.   usage: python tools/ncu_source_summary.py file.csv [top=40]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
def f(r, name):
    try: return float(r[col[name]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in body:
    for s in stalls: tot[s] += f(r, s)
ns = sum(f(r, "# Samples") for r in body); ni = sum(f(r, "Instructions Executed") for r in body)
nt = sum(f(r, "Thread Instructions Executed") for r in body)
print("instructions (SASS lines): %d   samples: %d   warp instr executed: %.4g   thread instr: %.4g   avg threads: %.2f" % (len(body), ns, ni, nt, nt / max(ni, 1)))
print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(ns, 1)) for k, v in tot.most_common(10)))
print("\n-- top by samples")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
    why = max(stalls, key=lambda s: f(r, s))
    print("%6d smp %5.2f%%  exec %10.0f thr %5.1f  %-12s %s" % (f(r, "# Samples"), 100 * f(r, "# Samples") / ns, f(r, "Instructions Executed"), f(r, "Avg. Threads Executed"), why[6:], r[col["Source"]].strip()[:90]))
print("\n-- cumulative executed warp instructions by opcode")
ops = collections.Counter()
for r in body:
    t = r[col["Source"]].strip().split()
    if not t: continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    ops[op.split(".")[0]] += f(r, "Instructions Executed")
print(", ".join("%s %.1f%%" % (k, 100 * v / ni) for k, v in ops.most_common(25)))
