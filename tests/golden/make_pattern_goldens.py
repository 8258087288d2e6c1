"""The reference's regression goldens for pattern.c: the command lines of c/ch5/makefile:49-62 and the lines the
reference printed for them (c/ch5/output/pattern.test1-5: program output, a few lines of numbers each).

Run in the build container (needs /root/reference):  python tests/golden/make_pattern_goldens.py
Writes tests/golden/pattern_goldens.json, which the tests read on machines without the reference tree.
"""
import json
import os
import re

REF = "/root/reference/c/ch5"
out = {}
mk = open(os.path.join(REF, "makefile")).read()
for n in range(1, 6):
    m = re.search(r'testit\.sh pattern "([^"]*)" (\d+) %d\b' % n, mk)
    lines = open(os.path.join(REF, "output", "pattern.test%d" % n)).read().splitlines()
    out["pattern.test%d" % n] = {"options": m.group(1), "ranks": int(m.group(2)),
                                 "source": "c/ch5/output/pattern.test%d" % n, "lines": lines}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "pattern_goldens.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
