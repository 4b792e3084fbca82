#!/bin/bash
# GPU call: A/B benches with env knobs: tools/gpu_run6.sh TAG name1:ENV=..,ENV=.. name2:...
mkdir -p gpurun_out
T=$1; shift
(time timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_vae_gpu.py -m gpu -q --tb=short -x -k "own_output or vae") > gpurun_out/${T}_unit.log 2>&1
tail -2 gpurun_out/${T}_unit.log
if grep -q "failed" gpurun_out/${T}_unit.log; then tail -40 gpurun_out/${T}_unit.log; exit 1; fi
names=""
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; envs=${envs//,/ }
  names="$names $name"
  env $envs timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/${T}_prof_$name.json > gpurun_out/${T}_bench_$name.log 2>&1
  grep '^{' gpurun_out/${T}_bench_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
done
python - $T $names <<'PY'
import json, sys
T = sys.argv[1]
for n in sys.argv[2:]:
    f = "gpurun_out/%s_prof_%s.json" % (T, n)
    try:
        d=json.load(open(f))
    except Exception as e:
        print(n, "missing", e); continue
    pk=d["per_kernel_family"]; tot=sum(v["ms"] for v in pk.values())
    print(n, "ms/step", round(d["ms_per_step_unprofiled"],1), "serial sum", round(tot,1))
    for k,v in sorted(pk.items(), key=lambda kv:-kv[1]["ms"])[:16]:
        print("   %8.2f ms %5d  %s  %.0f TF/s %.0f GB/s" % (v["ms"], v["launches"], k, v["flops_alg"]/max(v["ms"],1e-9)/1e9, v["bytes"]/max(v["ms"],1e-9)/1e6))
PY
if [ -n "$FULL" ]; then (time timeout 900 python -m pytest tests -m gpu -q --tb=short -x) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log; fi
exit 0
