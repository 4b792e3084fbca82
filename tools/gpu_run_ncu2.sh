#!/bin/bash
# ncu evidence for the state with GroupNorm in the convolution epilogues (round 2, second half): launch list of one timed
# bench step, per-kernel metrics of one VAE chunk, --set full of the convolutions that normalise their own output.
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,sm__cycles_elapsed.avg,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1400 --csv \
  --log-file gpurun_out/r2b_launches_step.csv python bench.py --steps 1 --warmup 2 --no-graph --no-cpu-baseline --no-extra --prof-steps 0 --ncu-range > gpurun_out/r2b_ncu_launches.log 2>&1
PART=vae timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -c 110 --csv --log-file gpurun_out/r2b_vae_metrics.csv python tools/gpu_ncu_target.py > gpurun_out/r2b_ncu_vae.log 2>&1
PART=convnorm timeout 300 ncu --metrics $M --clock-control none --profile-from-start off -c 6 --csv --log-file gpurun_out/r2b_convnorm_metrics.csv python tools/gpu_ncu_target.py > gpurun_out/r2b_ncu_convnorm.log 2>&1
PART=convnorm timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sw_kernel -c 3 -f -o gpurun_out/r2b_full_convnorm python tools/gpu_ncu_target.py > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/r2b_full_convnorm.ncu-rep > gpurun_out/r2b_full_convnorm.ncu.txt 2>&1
python tools/summarize_launches.py gpurun_out/r2b_launches_step.csv > gpurun_out/r2b_launches_summary_step.txt 2>&1
python tools/summarize_metrics.py gpurun_out/r2b_vae_metrics.csv > gpurun_out/r2b_vae_kernel_metrics.txt 2>&1
python tools/summarize_metrics.py gpurun_out/r2b_convnorm_metrics.csv > gpurun_out/r2b_convnorm_kernel_metrics.txt 2>&1
du -sh gpurun_out; tail -30 gpurun_out/r2b_convnorm_kernel_metrics.txt; tail -15 gpurun_out/r2b_launches_summary_step.txt
