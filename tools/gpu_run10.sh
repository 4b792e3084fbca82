#!/bin/bash
# GPU call: kernel + VAE tests on the shipped library, then config-3 A/B against an experiment / previous build (RGM_LIB)
mkdir -p gpurun_out
T=$1; ALT=$2
(timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_vae_gpu.py tests/test_flagship_gpu.py -m gpu -q --tb=short -x -k "dual or own_output or vae or flagship") 2>&1 | grep "VAE decode\|decoded roll\|margin\|passed\|failed" | cut -c1-200
for rep in 1 2 3; do
  for lib in "" "$ALT"; do
    name=$([ -z "$lib" ] && echo new || echo prev)
    RGM_LIB=$lib timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-steps 0 > gpurun_out/${T}_bench_${name}_$rep.log 2>&1
    grep '^{' gpurun_out/${T}_bench_${name}_$rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name $rep', round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"
  done
done
exit 0
