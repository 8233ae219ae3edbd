import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[2] if len(sys.argv)>2 else "", "n_gpus", d["n_gpus"], "ms/step %.2f"%d["ms_per_step"], "value %.3e"%d["value"], "step frac %.3f"%r["step"]["frac"], {k:round(v,2) for k,v in r["kernels_ms_per_step"].items()}, "elbo", d["elbo_trace_tail"])
