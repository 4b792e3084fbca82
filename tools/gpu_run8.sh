#!/bin/bash
# GPU call: A/B of an experiment build of the library (RGM_LIB) against the shipped one: VAE parity, then config-3 benches
mkdir -p gpurun_out
T=$1; ALT=$2
for lib in "" "$ALT"; do
  name=$([ -z "$lib" ] && echo base || echo alt)
  RGM_LIB=$lib timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_flagship_gpu.py -m gpu -q --tb=short 2>&1 | grep "rel-L2\|fraction\|ratio\|passed\|failed" | sed "s/^/$name /" | cut -c1-200
done
for rep in 1 2; do
  for lib in "" "$ALT"; do
    name=$([ -z "$lib" ] && echo base || echo alt)
    RGM_LIB=$lib timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-steps 0 > gpurun_out/${T}_bench_${name}_$rep.log 2>&1
    grep '^{' gpurun_out/${T}_bench_${name}_$rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name $rep', round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"
  done
done
exit 0
