"""Aggregate an ncu report's SASS-level samples / executed instructions by CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep cubin-name kernel-mangled-substring [top]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import Counter

rep, cubin_name, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all /root/repo/snap_b200/libsnapb200.so > /dev/null", shell=True, check=True)
dis = subprocess.run(f"nvdisasm -g -c {tmp}/{cubin_name}.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.split("\n")
off2line, cur, infun = {}, None, False
for ln in dis:
    if ln.startswith(".text.") and ksub in ln and ln.endswith(":"):
        infun = True
        continue
    if infun:
        if ln.startswith("//--------------------- ."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hi = next(i for i, r in enumerate(rows) if "Address" in r)
hdr = rows[hi]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[hi + 1][ia], 16)
ex, sm = Counter(), Counter()
for r in rows[hi + 1:]:
    if len(r) <= iexec:
        continue
    key = off2line.get(int(r[ia], 16) - base)
    ex[key] += int(r[iexec] or 0)
    sm[key] += int(r[isamp] or 0)
te, ts = sum(ex.values()), sum(sm.values())
src = {}
print(f"total warp-instructions {te}, samples {ts}")
for key, n in sm.most_common(top):
    f, l = key if key else ("?", 0)
    if f not in src:
        try:
            src[f] = open("/root/repo/snap_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = None
    text = src[f][l - 1].strip()[:100] if src[f] and l > 0 else ""
    print(f"{sm[key] / ts * 100:5.1f}% smp {ex[key] / te * 100:5.1f}% ex  {f}:{l}  {text}")
