"""Pretty-print a bench.py JSON line (kernel table + headline)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "latency_bs1", "cpu_baseline")})
print(d.get("roofline"))
for k in d.get("kernels") or []:
    print(f"{k['kernel']:26s} n={k['launches']:5d} ms={k['ms']:8.3f} share={k['share']:.3f} GB/s={k['gbs']:8.1f} TF={k['tflops']:6.2f}")
