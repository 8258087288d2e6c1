"""Extract the numbers the reference's own regression goldens pin for the fish path.

Run in the build container (needs /root/reference):  python tests/golden/make_fish_goldens.py
Reads c/ch6/output/fish.test1-8 and the matching command lines in c/ch6/makefile:11-33 and
writes tests/golden/fish_goldens.json.  Only the printed numbers are kept (no source text).
"""
import json
import os
import re

REF = "/root/reference/c/ch6"
out = {}
mk = open(os.path.join(REF, "makefile")).read()
for n in range(1, 9):
    m = re.search(r'testit\.sh fish "([^"]*)" (\d+) %d\n' % n, mk)
    txt = open(os.path.join(REF, "output", "fish.test%d" % n)).read()
    e = {"options": m.group(1), "ranks": int(m.group(2)), "source": "c/ch6/output/fish.test%d" % n}
    mm = re.search(r"iterations (\d+)", txt)
    if mm:
        e["ksp_its"] = int(mm.group(1))
    mm = re.search(r"0 SNES Function norm ([0-9.eE+-]+)", txt)
    if mm:
        e["snes_fnorm0"] = mm.group(1)
    mm = re.search(r"problem (\w+) on (.*) grid:", txt)
    e["problem"], e["gridstr"] = mm.group(1), mm.group(2)
    mm = re.search(r"_inf = ([0-9.eE+-]+), \|u-uexact\|_h = ([0-9.eE+-]+)", txt)
    e["errinf"], e["err2h"] = mm.group(1), mm.group(2)
    e["symmetric_lines"] = len(re.findall(r"Matrix is symmetric", txt))
    out["fish.test%d" % n] = e
json.dump(out, open(os.path.join(os.path.dirname(__file__), "fish_goldens.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
