FUZZ_CASES=150 FUZZ_SEED=41 timeout 400 python tools/fuzz_ldvm.py > gpurun_out/s3p_fuzz_ldvm.txt 2>&1; echo "fuzz rc=$?"; grep -c "skipped" gpurun_out/s3p_fuzz_ldvm.txt; tail -2 gpurun_out/s3p_fuzz_ldvm.txt; grep "above tolerance\|worst" gpurun_out/s3p_fuzz_ldvm.txt | head -4
FUZZ_CASES=120 FUZZ_SEED=42 FUZZ_MULTI=0 UNSFLOW_CLUSTER_SIZE=8 timeout 300 python tools/fuzz_ldvm.py > gpurun_out/s3p_fuzz_ldvm_c8.txt 2>&1; echo "fuzz c8 rc=$?"; grep "above tolerance\|worst" gpurun_out/s3p_fuzz_ldvm_c8.txt | head -3
python -m pytest tests -m gpu -x -q > gpurun_out/s3p_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3p_gputests.log; tail -3 gpurun_out/s3p_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/s3p_bench_1gpu.json 2> gpurun_out/s3p_bench_1gpu.err; echo "bench rc=$?"
