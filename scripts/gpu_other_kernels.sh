# duration + DRAM bytes of the non-K3 kernels on the bench workload.  Two passes: the construction/preprocessing kernels without the
# solver (K4 launches ~900 times per batch: profile it separately with -c), then a few K4 / re-plan launches.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/other_kernels_a.csv -k regex:"k1_|k2_|k5_|k6_|merge_maps|build_duo|permute_duos|transform_cells" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reg-steps 0 --pre-scans 2 > /dev/null 2> gpurun_out/other_kernels.err; tail -2 gpurun_out/other_kernels.err
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/other_kernels_b.csv -k regex:"k4_" -c 40 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reg-steps 1 --reg-streams 0 --pre-scans 0 > /dev/null 2>> gpurun_out/other_kernels.err
python scripts/ncu_kernel_table.py gpurun_out/other_kernels_a.csv gpurun_out/other_kernels_b.csv
