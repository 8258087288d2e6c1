"""The reference's regression goldens for minimal.c: the command lines of c/ch7/makefile:15-25 and the lines the
reference printed for them (c/ch7/output/minimal.test1-4: program output, a few lines of numbers each).

Run in the build container (needs /root/reference):  python tests/golden/make_minimal_goldens.py
Writes tests/golden/minimal_goldens.json, which the tests read on machines without the reference tree.
"""
import json
import os
import re

REF = "/root/reference/c/ch7"
out = {}
mk = open(os.path.join(REF, "makefile")).read()
for n in range(1, 5):
    m = re.search(r'testit\.sh minimal "([^"]*)" (\d+) %d\n' % n, mk)
    lines = open(os.path.join(REF, "output", "minimal.test%d" % n)).read().splitlines()
    out["minimal.test%d" % n] = {"options": m.group(1), "ranks": int(m.group(2)),
                                 "source": "c/ch7/output/minimal.test%d" % n, "lines": lines}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "minimal_goldens.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
