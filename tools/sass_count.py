"""Count SASS opcodes per kernel:  python tools/sass_count.py file.sass [name-substring]   (input: cuobjdump -sass output)."""
import collections
import re
import sys

fn, counts = None, collections.defaultdict(collections.Counter)
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        counts[fn][m.group(1)] += 1
sel = sys.argv[2] if len(sys.argv) > 2 else ""
for f, c in counts.items():
    if sel in f:
        fp = c["DFMA"] + c["DADD"] + c["DMUL"]
        print(f, "total", sum(c.values()), "fp64", fp, {k: v for k, v in c.most_common(14)})
