#!/bin/bash
# GPU call: the round's final records -- GPU test-suite, smoke, bench lines for configs 3 / 2 / 5, serial step profile
mkdir -p gpurun_out
T=${1:-r2z}
(time timeout 900 python -m pytest tests -m gpu -q --tb=short -x) > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.log
(time timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --prof-out gpurun_out/${T}_step_profile_serial.json > gpurun_out/${T}_bench_c3.log 2>&1; grep '^{' gpurun_out/${T}_bench_c3.log | cut -c1-400
timeout 300 python bench.py --config c2 --no-cpu-baseline > gpurun_out/${T}_bench_c2.log 2>&1; grep '^{' gpurun_out/${T}_bench_c2.log | cut -c1-300
timeout 300 python bench.py --config c5 --no-cpu-baseline > gpurun_out/${T}_bench_c5.log 2>&1; grep '^{' gpurun_out/${T}_bench_c5.log | cut -c1-300
