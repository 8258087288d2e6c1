"""The reference's regression goldens for heat.c: the command lines of c/ch5/makefile:43-47 and the lines the reference
printed for them (c/ch5/output/heat.test1-2: the banner and the -ts_monitor lines; test2's adaptive step sequence is what
pins [PETSc]'s default explicit scheme RK3bs and its step controller).

Run in the build container (needs /root/reference):  python tests/golden/make_heat_goldens.py
Writes tests/golden/heat_goldens.json, which the tests read on machines without the reference tree.
"""
import json
import os
import re

REF = "/root/reference/c/ch5"
out = {}
mk = open(os.path.join(REF, "makefile")).read()
for n in range(1, 3):
    m = re.search(r'testit\.sh heat "([^"]*)" (\d+) %d\b' % n, mk)
    lines = open(os.path.join(REF, "output", "heat.test%d" % n)).read().splitlines()
    out["heat.test%d" % n] = {"options": m.group(1), "ranks": int(m.group(2)), "source": "c/ch5/output/heat.test%d" % n,
                              "lines": lines}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "heat_goldens.json"), "w"), indent=1)
print({k: (v["options"], v["ranks"], len(v["lines"])) for k, v in out.items()})
