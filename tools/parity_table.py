"""gpurun_out/parity_table.jsonl (written by the GPU tests through tests/util.record_parity) -> a markdown table of the
measured errors per block, with the tolerance each test enforces:  python tools/parity_table.py > profiles/r02_parity_table.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_table.jsonl")
rows = {}
for ln in open(src):
    r = json.loads(ln)
    rows[(r["block"], r["what"])] = r          # the last run wins
print("# Measured parity errors (B200, `pytest -m gpu`), CUDA path vs the CPU oracle on identical inputs\n")
print("Relative L2 unless stated; every row is asserted by a test at the tolerance shown (<= 1.5 x the recorded error or the")
print("north_star bound of 1e-3, whichever the test states).  Integer / boolean outputs (visibility, tap indices, top-k view")
print("indices, valid planes, template validity, overlap counts, -inf masks) are compared bit-exactly and do not appear here.\n")
print("| block | quantity | measured | tolerance in the test |")
print("|---|---|---|---|")
for (block, what), r in sorted(rows.items()):
    print(f"| {block} | {what} | {r['err']:.3e} | {r['tol']:.1e} |")
