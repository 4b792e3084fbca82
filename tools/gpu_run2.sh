#!/bin/bash
# GPU call: full GPU test-suite, LayerNorm kernel variants on config 2, config 3 with a serial profile
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --tb=short -x) > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2b_pytest_gpu.log
for v in "RGM_LN_ONE_ROW=1" "RGM_LN_OCC=1" "RGM_LN_OCC=2"; do
  env $v timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --prof-out gpurun_out/r2b_prof_c2_$v.json > gpurun_out/r2b_bench_c2_$v.log 2>&1
done
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/r2b_step_profile.json > gpurun_out/r2b_bench_c3.log 2>&1
RGM_CONV_GN=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/r2b_step_profile_nofuse.json > gpurun_out/r2b_bench_c3_nofuse.log 2>&1
python - <<EOF
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_prof_c2_*.json")):
    d=json.load(open(f)); pk=d["per_kernel_family"]
    print(f, "ms/step", round(d["ms_per_step_unprofiled"],2), "ln_modulate ms", round(pk["ln_modulate"]["ms"],3), "GB/s", round(pk["ln_modulate"]["bytes"]/pk["ln_modulate"]["ms"]/1e6))
d=json.load(open("gpurun_out/r2b_step_profile.json")); print("c3 ms/step", d["ms_per_step_unprofiled"])
EOF
python - <<XEOF
import json
for f in ("gpurun_out/r2b_step_profile.json", "gpurun_out/r2b_step_profile_nofuse.json"):
    d=json.load(open(f)); pk=d["per_kernel_family"]; tot=sum(v["ms"] for v in pk.values())
    print(f, "ms/step", round(d["ms_per_step_unprofiled"],1), "serial sum", round(tot,1))
    for n,v in sorted(pk.items(), key=lambda kv:-kv[1]["ms"])[:14]:
        print("   %8.2f ms %5d  %s  %.0f TF/s" % (v["ms"], v["launches"], n, v["flops_alg"]/max(v["ms"],1e-9)/1e9))
XEOF
