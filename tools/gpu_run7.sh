#!/bin/bash
# GPU call: DiT A/B on config 2 (unguided p_sample, B=256) and config 3 with env knobs: TAG name:ENV=..,ENV=.. ...
mkdir -p gpurun_out
T=$1; shift
(time timeout 300 python -m pytest tests/test_dit_gpu.py tests/test_full_size_gpu.py -m gpu -q --tb=short -x) > gpurun_out/${T}_unit.log 2>&1
tail -2 gpurun_out/${T}_unit.log
if grep -q "failed" gpurun_out/${T}_unit.log; then tail -40 gpurun_out/${T}_unit.log; exit 1; fi
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; envs=${envs//,/ }
  for c in c2 c3; do
    env $envs timeout 300 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/${T}_prof_${name}_$c.json > gpurun_out/${T}_bench_${name}_$c.log 2>&1
    grep '^{' gpurun_out/${T}_bench_${name}_$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name $c', round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
  done
done
python - $T "$@" <<'PY'
import json, sys
T = sys.argv[1]
for spec in sys.argv[2:]:
    n = spec.split(":")[0]
    d = json.load(open("gpurun_out/%s_prof_%s_c2.json" % (T, n)))
    pk = d["per_kernel_family"]
    print(n, "c2 ms/step", round(d["ms_per_step_unprofiled"], 2), "serial sum", round(sum(v["ms"] for v in pk.values()), 2))
    for k, v in sorted(pk.items(), key=lambda kv: -kv[1]["ms"])[:9]:
        print("   %8.2f ms %5d  %s  %.0f TF/s" % (v["ms"], v["launches"], k, v["flops_alg"] / max(v["ms"], 1e-9) / 1e9))
PY
exit 0
