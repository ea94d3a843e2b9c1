python -m pytest tests -m gpu -x -q > gpurun_out/s3r_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3r_gputests.log; tail -3 gpurun_out/s3r_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/s3r_bench_1gpu.json 2> gpurun_out/s3r_bench_1gpu.err; echo "bench rc=$?"
FUZZ_CASES=150 FUZZ_SEED=51 FUZZ_MULTI=0 timeout 300 python tools/fuzz_ldvm.py > gpurun_out/s3r_fuzz_ldvm.txt 2>&1; echo "fuzz rc=$?"; grep "above tolerance" gpurun_out/s3r_fuzz_ldvm.txt
GROW_STEPS=1500 GROW_WINDOW=250 UNSFLOW_CLUSTER_PROFILE=1 python tools/ldvm_bench.py > gpurun_out/s3r_ldvm_bench.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cluster_steps -s 4 -c 1 -f -o gpurun_out/s3r_cluster python tools/c1_run.py > gpurun_out/s3r_ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
