# batched raw-scan filter: bench preprocess leg (wall clock) + ncu duration / DRAM bytes of the K6 kernels (single scan and 64-scan batch)
mkdir -p gpurun_out
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --reg-steps 0 --pre-scans 8 > gpurun_out/bench_k6b.json 2> gpurun_out/bench_k6b.err; tail -3 gpurun_out/bench_k6b.err
python -c "import json; d=json.load(open('gpurun_out/bench_k6b.json')); print(json.dumps(d['preprocess'], indent=1))"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/k6_batch.csv -k regex:"k6_" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reg-steps 0 --pre-scans 2 > /dev/null 2> gpurun_out/k6_batch.err; tail -2 gpurun_out/k6_batch.err
python scripts/ncu_kernel_table.py gpurun_out/k6_batch.csv
