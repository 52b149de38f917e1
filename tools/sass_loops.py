"""Static view of a kernel's SASS: every backward branch (= loop), its body size and opcode mix.
usage: python tools/sass_loops.py <lib.so> <kernel-name-substring>"""
import collections, re, subprocess, sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print("==", name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
            j = addr[int(m.group(1), 16)]
            body = ins[j:i + 1]
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            print("  loop 0x%04x..0x%04x: %d instr  %s" % (ins[j][0], a, len(body), " ".join("%s:%d" % kv for kv in ops.most_common(14))))
