#!/bin/bash
# GPU call: GPU test-suite; config 3 serial profiles: fused / unfused GroupNorm conv, parity-major / parity-fastest
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --tb=short -x) > gpurun_out/r2c_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2c_pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/r2c_prof_$name.json > gpurun_out/r2c_bench_$name.log 2>&1
}
run fused RGM_CONV_GN=1
run nofuse RGM_CONV_GN=0
run nofuse_parmajor RGM_CONV_GN=0 RGM_PAR_FAST=0
python - <<EOF
import json
for n in ("fused", "nofuse", "nofuse_parmajor"):
    f = "gpurun_out/r2c_prof_%s.json" % n
    try:
        d=json.load(open(f))
    except Exception as e:
        print(n, "missing", e); continue
    pk=d["per_kernel_family"]; tot=sum(v["ms"] for v in pk.values())
    print(n, "ms/step", round(d["ms_per_step_unprofiled"],1), "serial sum", round(tot,1))
    for k,v in sorted(pk.items(), key=lambda kv:-kv[1]["ms"])[:18]:
        print("   %8.2f ms %5d  %s  %.0f TF/s %.0f GB/s" % (v["ms"], v["launches"], k, v["flops_alg"]/max(v["ms"],1e-9)/1e9, v["bytes"]/max(v["ms"],1e-9)/1e6))
EOF
