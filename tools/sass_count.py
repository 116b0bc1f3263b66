#!/usr/bin/env python
"""Static SASS opcode counts per kernel: cuobjdump -sass lib.so > all.sass; tools/sass_count.py all.sass SUBSTR..."""
import collections, re, sys
txt = open(sys.argv[1]).read()
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if not any(s in name for s in sys.argv[2:]):
        continue
    ops = collections.Counter()
    for line in f.split('\n'):
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            ops[m.group(2).split('.')[0]] += 1
    print(name[:70], sum(ops.values()), ops.most_common(16))
