"""Pretty-print the chain probes of a bench.py JSON line (tame_debug_probes layout, per probe warp t=0 / t=T-1:
[0] start ns, [1] chain wait for the NEXT node's inputs, [2] end ns, [3] helper unit-wait, [4] helper hand-wait, [5] helper chain-wait,
[6] cells, [7] chain input-wait; waits in SM cycles)."""
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]; p = d["chain_probes"]
def show(name, q):
    start, wn, end, wu, wh, wc, nodes, wi = q[:8]
    nodes = max(nodes, 1)
    print(f"{name}: end +{(end-start)/1e3:.0f}us | per node {(end-start)/1e3/nodes:.3f}us | chain waits for the next node's inputs {wn/1.9e3/nodes:.3f} | "
          f"helper waits: unit {wu/1.9e3/nodes:.3f} hand {wh/1.9e3/nodes:.3f} chain {wc/1.9e3/nodes:.3f} | chain waits for inputs {wi/1.9e3/nodes:.3f} (us per node @1.9GHz)")
print(d["ms_per_step"], d["roofline"].get("kernels_ms_per_step"))
show("t=0  ", p[0:8]); show("t=T-1", p[8:16])
