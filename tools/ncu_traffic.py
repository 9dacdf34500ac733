"""Regenerate the DRAM-traffic evidence of the benchmarked build (run ON the GPU box, under gpurun, ONE GPU):

    python tools/ncu_traffic.py [--tag r02]        # writes gpurun_out/traffic.json + gpurun_out/<tag>_*_ncu_full.txt

`ncu --set full` captures (cold caches: ncu flushes L2 before every replay, so the numbers are the compulsory traffic) of
  * the fused lift            (tools/lift_one.py, one launch = 8 tiles),
  * the exhaustive correlation (tools/xcorr_one.py, G=128 R=36, one example),
  * every conv GEMM / GroupNorm-apply launch of ONE eager bench step (bench.py --profile-step, 8 tiles),
read back with `ncu -i ... --page raw --csv`.  Every entry carries the hash of the kernel source it measured
(bench.KERNEL_SOURCES); bench.py refuses an entry whose hash differs from the tree it runs in.  Copy
gpurun_out/traffic.json to profiles/traffic.json and the *_ncu_full.txt summaries to profiles/ afterwards.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (KERNEL_SOURCES / source_hash only; no CUDA work at import)

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="r02")
ap.add_argument("--skip-step", action="store_true", help="skip the whole-step capture (GEMM / GroupNorm kernels)")
args = ap.parse_args()
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


REPS = os.environ.get("SNAPB200_NCU_REP_DIR", "/tmp/snapb200_ncu")   # raw reports stay OFF gpurun_out/ (64 MiB merge limit)
os.makedirs(REPS, exist_ok=True)
LIGHT = ("dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,"
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,"
         "sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")


def capture(name, regex, cmd, count, light=False, skip=0):
    rep = os.path.join(REPS, f"{args.tag}_{name}")
    what = ["--metrics", LIGHT] if light else ["--set", "full", "--import-source", "on"]
    full = ["ncu", *what, "--clock-control", "none", "-k", f"regex:{regex}", "--profile-from-start", "off", "-s", str(skip),
            "-c", str(count), "-f", "-o", rep] + cmd
    r = subprocess.run(full, capture_output=True, text=True, cwd=ROOT)
    if r.returncode != 0:
        print(f"ncu failed for {name}: {r.stdout[-400:]} {r.stderr[-400:]}", file=sys.stderr)
        return None
    return rep + ".ncu-rep"


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    return hdr, rows[2:]


def col(hdr, name):
    return hdr.index(name)


def summarize(rep, txt_name):
    s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    with open(os.path.join(OUT, txt_name), "w") as f:
        f.write(s)


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def kernel_stats(rep):
    """[(kernel name, dram bytes, duration us, tensor pipe % of elapsed)] per profiled launch."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    it = hdr.index("gpu__time_duration.sum")
    ip = hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") if \
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in hdr else None
    out = []
    for v in vals:
        if len(v) <= max(ir, iw, it):
            continue
        dur = float(v[it]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[it], 1.0)
        out.append((v[ik], to_bytes(v[ir], units[ir]) + to_bytes(v[iw], units[iw]), dur,
                    float(v[ip]) if ip is not None and v[ip] else None))
    return out


traffic = {}
py = sys.executable
rep = capture("lift", "lift_fused2", [py, "tools/lift_one.py", "--batch", "8", "--reps", "1", "--kernels", "v2"], 1)
if rep:
    name, nbytes, dur, tens = kernel_stats(rep)[0]
    traffic["lift"] = {"kernel": name, "dram_bytes_per_launch": nbytes, "units_per_launch": 8, "duration_us": dur,
                       "tensor_pipe_pct_of_elapsed": tens, "src_sha": bench.source_hash("lift"),
                       "source": f"profiles/{args.tag}_lift_ncu_full.txt (ncu --set full, cold L2, one launch = 8 cfg2 tiles)"}
    summarize(rep, f"{args.tag}_lift_ncu_full.txt")
rep = capture("xcorr", "xcorr_rows_kernel", [py, "tools/xcorr_one.py", "--reps", "1"], 1)
if rep:
    name, nbytes, dur, tens = kernel_stats(rep)[0]
    traffic["xcorr"] = {"kernel": name, "dram_bytes_per_launch": nbytes, "units_per_launch": 1, "duration_us": dur,
                        "tensor_pipe_pct_of_elapsed": tens, "src_sha": bench.source_hash("xcorr"),
                        "source": f"profiles/{args.tag}_xcorr_ncu_full.txt (ncu --set full, cold L2, G=128 R=36, one example)"}
    summarize(rep, f"{args.tag}_xcorr_ncu_full.txt")
if not args.skip_step:
    # every GEMM / GroupNorm launch of one eager step with a light metric set (DRAM bytes, duration, tensor pipe, issue
    # slots), then `--set full` summaries of the three instantiations that dominate the step
    rep = capture("step", "gemm_tc_kernel|gn_apply_kernel|conv3x3_halo_kernel", [py, "bench.py", "--profile-step", "--no-cpu-baseline"], 400, light=True)
    for nm, rx, skip in (("gemm_128_64_conv", "gemm_tc_kernel<128, 64, 0, 1>", 3), ("gemm_256_64", "gemm_tc_kernel<256, 64, 0, 0>", 2),
                         ("gn_apply_dense", "gn_apply_kernel<0, 0, 0>", 2), ("conv_gn_tgn1", "gemm_tc_kernel<64, 64, 3, 1>", 1),
                         ("conv3x3_halo", "conv3x3_halo_kernel<64, 1, 1>", 1)):
        r2 = capture(nm, rx.replace("(", ".").replace("<", ".").replace(">", ".").replace(", ", ".."), [py, "bench.py", "--profile-step", "--no-cpu-baseline"], 1, skip=skip)
        if r2:
            summarize(r2, f"{args.tag}_{nm}_ncu_full.txt")
    if rep:
        stats = kernel_stats(rep)
        for key, pat in (("gemm", "gemm_tc_kernel"), ("gn_apply", "gn_apply_kernel")):
            sel = [s for s in stats if pat in s[0] or (key == "gemm" and "conv3x3_halo_kernel" in s[0])]
            if not sel:
                continue
            by = {}
            for name, nbytes, dur, tens in sel:
                short = name.split("(")[0].replace("snapb200::", "").replace("void ", "")
                e = by.setdefault(short, {"launches": 0, "dram_bytes": 0.0, "duration_us": 0.0, "tensor_pct_x_us": 0.0})
                e["launches"] += 1
                e["dram_bytes"] += nbytes
                e["duration_us"] += dur
                e["tensor_pct_x_us"] += (tens or 0.0) * dur
            for e in by.values():
                e["dram_gbs"] = e["dram_bytes"] / (e["duration_us"] * 1e-6) / 1e9
                e["tensor_pipe_pct_of_elapsed"] = e.pop("tensor_pct_x_us") / e["duration_us"]
            traffic[key] = {"per_step_of_8_tiles": {"launches": len(sel), "dram_bytes": sum(s[1] for s in sel),
                                                    "duration_us": sum(s[2] for s in sel)},
                            "by_instantiation": by, "src_sha": bench.source_hash(key),
                            "source": f"profiles/{args.tag}_step_kernels.json (ncu, DRAM-bytes + duration + tensor-pipe metrics of every launch of one eager bench step, cold L2)"}
        with open(os.path.join(OUT, f"{args.tag}_step_kernels.json"), "w") as f:
            json.dump([{"kernel": s[0][:120], "dram_bytes": s[1], "duration_us": s[2], "tensor_pipe_pct": s[3]} for s in stats], f, indent=0)
with open(os.path.join(OUT, "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "by_instantiation"} for k, v in traffic.items()}, indent=1))
