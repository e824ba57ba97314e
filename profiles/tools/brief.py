"""one-line digest of a bench.py JSON line on stdin: step, e2e, stage times (used by A/B shell loops on the GPU box)"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print(tag, "step %.3f e2e %.3f |" % (d["ms_per_step"], d["e2e"]["ms_per_step"]), " ".join("%s %.3f" % (n, v["ms"]) for n, v in k.items()))
