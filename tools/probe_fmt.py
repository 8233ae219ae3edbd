import json,sys
d=json.load(open(sys.argv[1])); p=d["chain_probes"]; n=d["config"]["n"]
k0=p[7]
def show(name,q):
    start,first,end,wu,wh,gj,nodes=q[:7]
    print(f"{name}: start+{(start-k0)/1e3:.0f}us first-node+{(first-k0)/1e3:.0f}us end+{(end-k0)/1e3:.0f}us | per node {(end-first)/1e3/max(nodes,1):.2f}us | unit-wait {wu/1.9e3/max(nodes,1):.2f}us hand-wait {wh/1.9e3/max(nodes,1):.2f}us gj {gj/1.9e3/max(nodes,1):.2f}us (per node, @1.9GHz)")
print(d["ms_per_step"], d["roofline"]["kernels_ms_per_step"])
show("t=0  ",p[0:8])
if len(sys.argv) > 2: show("t=T-1",p[8:16])
