"""Aggregate the per-SASS-instruction samples of an ncu report by CUDA source line.
usage: ncu_lines.py <sass csv from `ncu -i rep --page source --csv`> <nvdisasm -g -c listing of the same function> [top]
(the SASS page carries no line numbers; nvdisasm's `//## File ..., line N` annotations do, and both list the function's
instructions in the same order, so they are joined by offset)."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
ins = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2]
base = int(ins[0][0], 16)
line_of = {}; cur = None
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for r in ins:
    off = int(r[0], 16) - base
    ln = line_of.get(off)
    s = int(r[ix["# Samples"]] or 0); tot += s
    a = agg[ln]; a["samples"] += s; a["inst"] += int(r[ix["Instructions Executed"]] or 0); a["n"] += 1
    for h in stalls: a[h] += int(r[ix[h]] or 0)
src = open("python-temporal-ame-svi_b200/csrc/tame_kernels.cuh").read().split("\n")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print("total samples", tot, "sass instructions", len(ins))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ", ".join(f"{h[6:]} {a[h]}" for h in sorted(stalls, key=lambda h: -a[h])[:3] if a[h])
    text = src[ln - 1].strip()[:90] if ln else "?"
    print(f"{str(ln):>5} {a['samples']:7d} {100*a['samples']/max(tot,1):5.1f}% sass={a['n']:4d} exec={a['inst']:9d} | {st:45s} | {text}")
if len(sys.argv) > 5:
    lo, hi = int(sys.argv[4]), int(sys.argv[5])
    c = collections.Counter(); n = 0; ex = 0
    for ln, a in agg.items():
        if ln is not None and lo <= ln < hi:
            for h in stalls: c[h] += a[h]
            n += a["samples"]; ex += a["inst"]
    print(f"lines [{lo},{hi}): samples {n} inst {ex}", {k[6:]: v for k, v in c.most_common(9)})
if len(sys.argv) > 6:
    per = float(sys.argv[6])
    print("by instructions executed (per cell = exec / %g):" % per)
    for ln, a in sorted(((l, a) for l, a in agg.items() if l is not None and lo <= l < hi), key=lambda kv: -kv[1]["inst"])[:45]:
        print(f"{ln:5d} exec/cell {a['inst']/per:7.1f} samples {a['samples']:5d} | {src[ln-1].strip()[:100]}")
