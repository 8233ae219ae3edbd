#!/usr/bin/env python
"""SASS instruction counts per kernel of libtame_b200.so (cuobjdump -sass): FP64 tensor-core (DMMA), FP64 FMA (DFMA),
asynchronous global->shared copies (LDGSTS = cp.async), TMA bulk copies (UBLKCP / UTMALDG), mbarrier (SYNCS), reciprocal
(MUFU.RCP64H).  Written to profiles/ so that claims about which units the kernels use are checkable.
    python tools/sass_counts.py > profiles/r02_sass_counts.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "python-temporal-ame-svi_b200", "tame_b200", "libtame_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
keys = ["DMMA", "DFMA", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "MUFU.RCP64H", "LDS", "STG", "total"]
for line in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(1); c = counts[kern]; c["total"] += 1
        for k in keys[:-1]:
            if op == k or op.startswith(k + "."): c[k] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.split("\n")
print("# SASS instruction counts per kernel (sm_100a cubins of libtame_b200.so; `python tools/sass_counts.py`)\n")
print("Static counts (instructions in the binary, not executed).  `DMMA` = `mma.sync.m8n8k4.f64`; `LDGSTS` = `cp.async`;")
print("`UBLKCP`/`UTMALDG` = TMA bulk copies: none -- the staging is per-thread `cp.async` (profiles/r01_summary.md, 'Things that did not work').\n")
print("| kernel | " + " | ".join(keys) + " |\n|---|" + "---|" * len(keys))
want = ("k_sweep", "k_chain", "k_contract", "k_llmse", "k_covblend", "k_cellterms", "k_align", "k_totals", "k_hab", "k_generate")
for (k, c), d in zip(counts.items(), dem):
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", d).replace("void ", "").replace("(int)", "").replace("(bool)", "").replace("(anonymous namespace)::", "")
    if not any(w in name for w in want): continue
    if re.search(r"<([1235-7])[,>]", name): continue          # keep the r = 4 and r = 8 instantiations
    print(f"| `{name}` | " + " | ".join(str(c[x]) for x in keys) + " |")
