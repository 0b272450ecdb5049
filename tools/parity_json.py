"""python tools/parity_json.py <run tag> : gpurun_out/parity.jsonl (written by tests/helpers.py:record during `pytest -m gpu` on the
B200 box) -> profiles/parity_r2.json, one entry per recorded case (the last record of a case wins)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "final"
cases = {}
for line in open(os.path.join(REPO, "gpurun_out", "parity.jsonl")):
    line = line.strip()
    if line:
        d = json.loads(line)
        cases[d["case"]] = d
out = {"source": f"tests -m gpu on B200, round 2 (run {tag}); one entry per recorded case: measured errors of the CUDA path against "
                 "the CPU oracle (cosine, rel-L2 per parameter tensor where listed)",
       "cases": list(cases.values())}
with open(os.path.join(REPO, "profiles", "parity_r2.json"), "w") as f:
    f.write("{\n" + json.dumps("source") + ": " + json.dumps(out["source"]) + ",\n\"cases\": [\n")
    f.write(",\n".join(json.dumps(c, indent=0) for c in out["cases"]))
    f.write("\n]\n}\n")
print(len(cases), "cases")
