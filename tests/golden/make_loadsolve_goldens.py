"""The reference's golden for loadsolve.c (c/ch2/makefile:34-36, c/ch2/output/loadsolve.test1): tri.c's 4 x 4 system
written to A.dat / b.dat, read back, solved, and viewed in ASCII.

Run in the build container (needs /root/reference):  python tests/golden/make_loadsolve_goldens.py
Writes tests/golden/loadsolve_goldens.json."""
import json
import os
import re

REF = "/root/reference/c/ch2"
mk = open(os.path.join(REF, "makefile")).read()
m = re.search(r'testit\.sh loadsolve "([^"]*)" (\d+) 1\b', mk)
out = {"loadsolve.test1": {"options": m.group(1), "ranks": int(m.group(2)), "source": "c/ch2/output/loadsolve.test1",
                           "lines": open(os.path.join(REF, "output", "loadsolve.test1")).read().split("\n")[:-1]}}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "loadsolve_goldens.json"), "w"), indent=1)
print(out["loadsolve.test1"]["options"], len(out["loadsolve.test1"]["lines"]))
