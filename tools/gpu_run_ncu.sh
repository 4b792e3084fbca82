#!/bin/bash
# ncu evidence, sized to come back through gpurun_out (<= 64 MiB): (1) the launch list of one timed bench step,
# (2) a metrics table of every kernel at bench-chunk shapes, (3) --set full of a few kernels, summarised on the box.
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,sm__cycles_elapsed.avg,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
# (1) launch list: one eager bench step bracketed by cudaProfilerStart/Stop (stops inside the VAE decode; see tools/summarize_launches.py)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1700 --csv \
  --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-extra --prof-steps 0 --ncu-range > gpurun_out/r2_ncu_launches.log 2>&1
# (2) metrics per kernel at bench-chunk shapes
PART=dit timeout 400 ncu --metrics $M --clock-control none --profile-from-start off -c 40 --csv --log-file gpurun_out/r2_dit_metrics.csv python tools/gpu_ncu_target.py > gpurun_out/r2_ncu_dit.log 2>&1
PART=vae timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -c 140 --csv --log-file gpurun_out/r2_vae_metrics.csv python tools/gpu_ncu_target.py > gpurun_out/r2_ncu_vae.log 2>&1
PART=convgn timeout 200 ncu --metrics $M --clock-control none --profile-from-start off -c 1 --csv --log-file gpurun_out/r2_convgn_metrics.csv python tools/gpu_ncu_target.py > gpurun_out/r2_ncu_convgn.log 2>&1
# (3) --set full: the dominant conv (256->256 @64x64, pair kernel), the 128-channel conv, attention, the fused conv
PART=vae timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_sw2_kernel -s 12 -c 2 -f -o gpurun_out/r2_full_sw2 python tools/gpu_ncu_target.py > /dev/null 2>&1
PART=vae timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_sw_kernel -s 2 -c 1 -f -o gpurun_out/r2_full_sw python tools/gpu_ncu_target.py > /dev/null 2>&1
PART=dit timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:attention_kernel -c 1 -f -o gpurun_out/r2_full_attention python tools/gpu_ncu_target.py > /dev/null 2>&1
PART=convgn timeout 300 ncu --set full --clock-control none --profile-from-start off -c 1 -f -o gpurun_out/r2_full_conv_gn python tools/gpu_ncu_target.py > /dev/null 2>&1
for f in sw2 sw attention conv_gn; do python tools/summarize_ncu.py gpurun_out/r2_full_$f.ncu-rep > gpurun_out/r2_full_$f.ncu.txt 2>&1; done
du -sh gpurun_out; ls -la gpurun_out | head -30
