#!/usr/bin/env python
"""Per-instruction stall samples of one profiled kernel (needs -lineinfo / --import-source on):
prints totals per opcode, the top-N stall lines and a coarse histogram along the program."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
extra = sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
print(rows[0][1][:150]); print("instructions", len(body), "samples", tot)
ops = collections.Counter(); cnt = collections.Counter()
for r in body:
    src = r[ix['Source']].split()
    op = (src[1] if src[0].startswith('@') else src[0]).split('.')[0]
    ops[op] += int(r[ix['# Samples']] or 0); cnt[op] += 1
print("by opcode:", [(k, cnt[k], v) for k, v in ops.most_common(12)])
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in body:
    for h in stall_cols:
        agg[h] += int(r[ix[h]] or 0)
print("by reason:", agg.most_common(10))
nb = 24; step = (len(body) + nb - 1) // nb
for b in range(nb):
    seg = body[b * step:(b + 1) * step]
    if not seg: break
    s = sum(int(r[ix['# Samples']] or 0) for r in seg)
    kinds = collections.Counter((r[ix['Source']].split()[1] if r[ix['Source']].startswith('@') else r[ix['Source']].split()[0]).split('.')[0] for r in seg)
    print(f"[{b*step:5d}-{(b+1)*step:5d}) {s:6d} {'#' * int(60 * s / max(tot,1))}  {dict(kinds.most_common(4))}")
for r in sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:topn]:
    st = {h[6:]: r[ix[h]] for h in stall_cols if r[ix[h]] not in ('0', '')}
    print(r[ix['Address']][-5:], r[ix['Source']][:70].ljust(70), r[ix['# Samples']], st)
