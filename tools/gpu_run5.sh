#!/bin/bash
# GPU call: GroupNorm-in-epilogue convolutions -- kernel test first (bounded), then the suite, then A/B benches
mkdir -p gpurun_out
T=${T:-r2g}
(time timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -x -k "own_output") > gpurun_out/${T}_unit.log 2>&1
tail -5 gpurun_out/${T}_unit.log
if ! grep -q " passed" gpurun_out/${T}_unit.log || grep -q "failed" gpurun_out/${T}_unit.log; then echo "UNIT TEST FAILED"; tail -40 gpurun_out/${T}_unit.log; exit 1; fi
(time timeout 300 python -m pytest tests/test_vae_gpu.py -m gpu -q --tb=short -x) > gpurun_out/${T}_vae.log 2>&1
tail -5 gpurun_out/${T}_vae.log
if grep -q "failed" gpurun_out/${T}_vae.log; then tail -40 gpurun_out/${T}_vae.log; exit 1; fi
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-out gpurun_out/${T}_prof_$name.json > gpurun_out/${T}_bench_$name.log 2>&1
  grep '^{' gpurun_out/${T}_bench_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
}
run epi_lanes2 RGM_GN_EPI=1 RGM_VAE_LANES=2
run epi_lanes1 RGM_GN_EPI=1 RGM_VAE_LANES=1
run sep_lanes2 RGM_GN_EPI=0 RGM_VAE_LANES=2
python - <<PY
import json
for n in ("epi_lanes2", "epi_lanes1", "sep_lanes2"):
    f = "gpurun_out/${T}_prof_%s.json" % n
    try:
        d=json.load(open(f))
    except Exception as e:
        print(n, "missing", e); continue
    pk=d["per_kernel_family"]; tot=sum(v["ms"] for v in pk.values())
    print(n, "ms/step", round(d["ms_per_step_unprofiled"],1), "serial sum", round(tot,1))
    for k,v in sorted(pk.items(), key=lambda kv:-kv[1]["ms"])[:16]:
        print("   %8.2f ms %5d  %s  %.0f TF/s %.0f GB/s" % (v["ms"], v["launches"], k, v["flops_alg"]/max(v["ms"],1e-9)/1e9, v["bytes"]/max(v["ms"],1e-9)/1e6))
PY
if [ -n "$FULL" ]; then (time timeout 900 python -m pytest tests -m gpu -q --tb=short -x) > gpurun_out/${T}_pytest_gpu.log 2>&1; fi
tail -3 gpurun_out/${T}_pytest_gpu.log
