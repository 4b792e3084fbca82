#!/bin/bash
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --tb=short -x) > gpurun_out/r2d_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2d_pytest_gpu.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/r2d_prof.json > gpurun_out/r2d_bench.log 2>&1
python - <<EOF
import json
d=json.load(open("gpurun_out/r2d_prof.json"))
pk=d["per_kernel_family"]; tot=sum(v["ms"] for v in pk.values())
print("ms/step", round(d["ms_per_step_unprofiled"],1), "serial sum", round(tot,1))
for k,v in sorted(pk.items(), key=lambda kv:-kv[1]["ms"])[:22]:
    print("   %8.2f ms %5d  %s  %.0f TF/s %.0f GB/s" % (v["ms"], v["launches"], k, v["flops_alg"]/max(v["ms"],1e-9)/1e9, v["bytes"]/max(v["ms"],1e-9)/1e6))
EOF
