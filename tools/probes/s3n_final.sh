python -m pytest tests -m gpu -x -q > gpurun_out/s3n_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3n_gputests.log; tail -3 gpurun_out/s3n_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s3n_smoke.log; tail -2 gpurun_out/s3n_smoke.log
python bench.py > gpurun_out/s3n_bench_1gpu.json 2> gpurun_out/s3n_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3n_bench_reference.json 2> gpurun_out/s3n_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/s3n_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s3n_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_induce_sym -s 2 -c 1 -f -o gpurun_out/s3n_sym8 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s3n_ncu_sym.log 2>&1; echo "ncu sym rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster_steps -s 4 -c 1 -f -o gpurun_out/s3n_cluster python tools/c1_run.py > gpurun_out/s3n_ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
ls -la gpurun_out/*.ncu-rep
