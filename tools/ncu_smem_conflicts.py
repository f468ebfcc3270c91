#!/usr/bin/env python
"""Per-instruction shared-memory wavefronts of one profiled kernel: which LDS / STS instructions replay
(L1 Wavefronts Shared vs Ideal), with the CUDA source line when the report has it.
Usage: python tools/ncu_smem_conflicts.py file.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
def num(r, k):
    try: return float(r[ix[k]] or 0)
    except ValueError: return 0.0
tot_w = sum(num(r, "L1 Wavefronts Shared") for r in body); tot_i = sum(num(r, "L1 Wavefronts Shared Ideal") for r in body)
print(rows[0][1][:140]); print(f"shared wavefronts {tot_w:.0f}, ideal {tot_i:.0f}, excess {tot_w - tot_i:.0f} ({100 * (tot_w - tot_i) / max(tot_w, 1):.1f} %)")
bad = sorted((r for r in body if num(r, "L1 Wavefronts Shared Excessive") > 0), key=lambda r: -num(r, "L1 Wavefronts Shared Excessive"))
for r in bad[:topn]:
    print(f"{r[ix['Address']][-5:]} {r[ix['Source']][:60]:60s} exec {num(r, 'Instructions Executed'):9.0f} waves {num(r, 'L1 Wavefronts Shared'):9.0f} "
          f"ideal {num(r, 'L1 Wavefronts Shared Ideal'):9.0f} excess {num(r, 'L1 Wavefronts Shared Excessive'):9.0f}")
