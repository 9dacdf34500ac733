"""Per-source-line stall-reason breakdown of an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_stalls.py report.ncu-rep cubin-name kernel-substring [top]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

rep, cubin_name, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
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
ia = hdr.index("Address")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {h: hdr.index(h) for h in reasons}
base = int(rows[hi + 1][ia], 16)
tot = Counter()
per_line = defaultdict(Counter)
for r in rows[hi + 1:]:
    if len(r) <= max(idx.values()):
        continue
    key = off2line.get(int(r[ia], 16) - base)
    for h in reasons:
        v = int(r[idx[h]] or 0)
        tot[h] += v
        per_line[h][key] += v
allsum = sum(tot.values())
print("stall reasons (all samples):", ", ".join(f"{h[6:]} {v / allsum * 100:.1f}%" for h, v in tot.most_common()))
for h, _ in tot.most_common(6):
    print(f"--- {h} ({tot[h] / allsum * 100:.1f}% of samples) top lines")
    for key, v in per_line[h].most_common(top):
        if v == 0:
            break
        f, l = key if key else ("?", 0)
        try:
            text = open("/root/repo/snap_b200/csrc/" + f).read().split("\n")[l - 1].strip()[:90]
        except Exception:
            text = ""
        print(f"   {v / tot[h] * 100:5.1f}%  {f}:{l}  {text}")
