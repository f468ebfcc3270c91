#!/usr/bin/env python
"""Turn the raw captures a GPU run left in gpurun_out/ into the committed artefacts under profiles/:
   python tools/make_profiles.py <tag>      (e.g. r1_d)
expects gpurun_out/{bench_full.json, bench_ref.json, launches.csv, prof_bench.ncu-rep, sweep_g.txt}
(tools/_run_bench.sh produces them)."""
import csv, json, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")

rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]; ix = {k: j for j, k in enumerate(h)}
with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e\n")
    f.write("# (per-launch times are cold-cache and serialised: compare shares, not absolutes)\nid,kernel,grid,block,duration_ns\n")
    tot = {}
    for r in rows[hdr + 1:]:
        if len(r) < len(h):
            continue
        name = r[ix["Kernel Name"]]
        m = re.search(r"(frame_kernel\w*)<b200fft::(\w+)<b200fft::Plan<b200fft::Arith(\w+), (?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+)", name)
        short = f"{m.group(1)}<{m.group(2)}<{m.group(3)},N={m.group(4)},T={m.group(5)},F={m.group(6)}>>" if m else name.split("(")[0][:70].replace(",", ";")
        f.write(f"{r[ix['ID']]},{short},{r[ix['Grid Size']].replace(', ', 'x')},{r[ix['Block Size']].replace(', ', 'x')},{r[ix['Metric Value']]}\n")
        tot[short] = tot.get(short, 0.0) + float(r[ix["Metric Value"]])
    s = sum(tot.values())
    f.write("# shares: " + "; ".join(f"{k} {100 * v / s:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:4]) + "\n")

summ = subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_summary.py"), os.path.join(G, "prof_bench.ncu-rep")], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_ncu_bench.txt"), "w").write(summ)
lines = [l.split() for l in summ.splitlines() if l.startswith("frame_kernel")]
cols = summ.splitlines()[1].split()
tr = {"_source": f"profiles/{tag}_ncu_bench.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, bench workload)"}
for l in lines:
    key = "rfft_fwd" if "RfftFwd" in l[0] else "rfft_inv"
    tr[key] = (float(l[cols.index("dram_rd_MB")]) + float(l[cols.index("dram_wr_MB")])) * 1e6
open(os.path.join(P, "traffic.json"), "w").write(json.dumps(tr, indent=1) + "\n")
for k, name in ((0, "rfft_fwd"), (1, "rfft_inv")):
    out = subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_hotspots.py"), os.path.join(G, "prof_bench.ncu-rep"), "25",
                          "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_hotspots_{name}.txt"), "w").write(f"# {name} (bench kernel), per-instruction stall samples\n" + "\n".join(l[:230] for l in out.splitlines()) + "\n")
for src, dst in (("bench_full.json", f"{tag}_bench.json"), ("bench_ref.json", f"{tag}_bench_reference_arm.json"), ("sweep_g.txt", f"{tag}_sweep.txt")):
    open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
for u in ("3_2048", "4_2048"):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(R, "cmsis-dsp_b200", "build", f"ku_{u}.o")], capture_output=True, text=True).stdout
    keep, on = [], False
    for l in sass.splitlines():
        if "Function :" in l:
            on = "frame_kernel_pipe" in l
        if on:
            keep.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/$", "", l))
    open(os.path.join(P, f"{tag}_sass_ku_{u}.pipe.sass"), "w").write("\n".join(keep) + "\n")
print("profiles written for", tag, tr)

# ---- the other kernel families (tools/gpu_round.sh: prof_fam_*.ncu-rep, sweep_rfix, config4)
fam = []
for name in ("mfcc", "cfft_f32", "cfft_q31", "cfft_q15", "rfftq31_fwd", "cfft_peak"):
    rep = os.path.join(G, f"prof_fam_{name}.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        fam.append(f"## {name}  (python tools/sweep.py --mib 256 --ops {name} --lens 1024 under ncu --set full --clock-control none)\n" + out)
        hs = subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_hotspots.py"), rep, "15"], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{tag}_ncu_hotspots_{name}.txt"), "w").write(f"# {name} N=1024, per-instruction stall samples\n" + "\n".join(l[:230] for l in hs.splitlines()) + "\n")
if fam:
    open(os.path.join(P, f"{tag}_ncu_families.txt"), "w").write("\n".join(fam))
for src, dst in (("sweep_rfix.txt", f"{tag}_sweep_rfft_fixed.txt"), ("config4_1gpu.json", f"{tag}_config4_1gpu.json")):
    if os.path.exists(os.path.join(G, src)):
        open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
for unit, pat, dst in (("mfcc", "mfcc_kernel_pipeILi512E", f"{tag}_sass_mfcc_pipe_1024.sass"), ("ku_1_1024", "Lb0ELb0ELb0ELb0E", f"{tag}_sass_ku_1_1024.sass")):
    obj = os.path.join(R, "cmsis-dsp_b200", "build", f"{unit}.o")
    if not os.path.exists(obj):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    keep, on = [], False
    for l in sass.splitlines():
        if "Function :" in l:
            on = pat in l
        if on:
            keep.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/$", "", l))
    if keep:
        open(os.path.join(P, dst), "w").write("\n".join(keep) + "\n")
print("family profiles written")
