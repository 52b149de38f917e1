"""Basic-block view of `ncu --page source --csv` (SASS rows with per-instruction counters): consecutive instructions with
the same execution count form a block; blocks are listed by their share of the executed warp instructions, with the
average number of active threads, the share of stall samples and the opcode mix.  With --sass N the N hottest blocks are
also printed instruction by instruction.
    usage: python tools/ncu_blocks.py file.csv [--table K] [--top N] [--sass N]"""
import collections, csv, sys

args = sys.argv[1:]
path = args[0]
def opt(name, default):
    return int(args[args.index(name) + 1]) if name in args else default
table, top, sass = opt("--table", -1), opt("--top", 30), opt("--sass", 0)
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[table]
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
end = his[his.index(hi) + 1] if his.index(hi) + 1 < len(his) else len(rows)
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
def f(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
blocks, cur = [], None
for r in body:
    e = f(r, "Instructions Executed")
    if cur and cur["e"] == e:
        cur["rows"].append(r)
    else:
        cur = {"e": e, "rows": [r]}; blocks.append(cur)
tot = sum(b["e"] * len(b["rows"]) for b in blocks); ts = sum(f(r, "# Samples") for r in body)
tthr = sum(f(r, "Thread Instructions Executed") for r in body)
print("%d SASS instructions, %.4g warp instructions executed, %.4g thread instructions (%.2f active threads per instruction), %d stall samples"
      % (len(body), tot, tthr, tthr / max(tot, 1), ts))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = collections.Counter()
for r in body:
    for s in stalls: st[s] += f(r, s)
print("stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(ts, 1)) for k, v in st.most_common(8)))
def op(o):
    t = o.split()
    return (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
order = sorted(blocks, key=lambda b: -b["e"] * len(b["rows"]))
print("\nblock (address range)   instr  executions  active thr  share of warp instr  share of samples  opcode mix")
for b in order[:top]:
    rs = b["rows"]; n = len(rs)
    thr = sum(f(r, "Thread Instructions Executed") for r in rs) / max(b["e"] * n, 1)
    smp = sum(f(r, "# Samples") for r in rs)
    ops = collections.Counter(op(r[col["Source"]].strip()) for r in rs)
    print("%s-%s  %4d  %10.0f  %5.1f  %6.2f%%  %6.2f%%  %s" % (rs[0][col["Address"]][-5:], rs[-1][col["Address"]][-5:], n, b["e"], thr, 100 * b["e"] * n / tot,
                                                          100 * smp / max(ts, 1), " ".join("%s:%d" % kv for kv in ops.most_common(9))))
for b in order[:sass]:
    rs = b["rows"]
    print("\n-- block %s-%s: %d instructions x %.0f executions" % (rs[0][col["Address"]][-5:], rs[-1][col["Address"]][-5:], len(rs), b["e"]))
    for r in rs:
        why = max(stalls, key=lambda s: f(r, s)) if f(r, "# Samples") else ""
        print("  %s  smp %5.0f %-14s %s" % (r[col["Address"]][-5:], f(r, "# Samples"), why[6:], r[col["Source"]].strip()[:110]))
