#!/bin/bash
# GPU call: A/B of an experiment build (RGM_LIB) on the DiT: parity tests, then config-2 and config-3 benches, twice
mkdir -p gpurun_out
T=$1; ALT=$2
for lib in "" "$ALT"; do
  name=$([ -z "$lib" ] && echo base || echo alt)
  RGM_LIB=$lib timeout 300 python -m pytest tests/test_dit_gpu.py tests/test_flagship_gpu.py tests/test_sampler_gpu.py -m gpu -q --tb=short 2>&1 | grep "DiT\|eps\|x_t after step 3\|chosen\|passed\|failed" | sed "s/^/$name /" | cut -c1-210
done
for rep in 1 2; do
  for lib in "" "$ALT"; do
    name=$([ -z "$lib" ] && echo base || echo alt)
    for c in c2 c3; do
      RGM_LIB=$lib timeout 300 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-extra --prof-steps 0 > gpurun_out/${T}_bench_${name}_${c}_$rep.log 2>&1
      grep '^{' gpurun_out/${T}_bench_${name}_${c}_$rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name $c $rep', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
    done
  done
done
exit 0
