python -m pytest tests -m gpu -x -q > gpurun_out/s3h_gputests_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3h_gputests_2gpu.log
tail -3 gpurun_out/s3h_gputests_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s3h_bench_2gpu.json 2> gpurun_out/s3h_bench_2gpu.err; echo "bench rc=$?"
UNSFLOW_STEP_TRACE=1 MG_WAKE=100000 MG_STEPS=10 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ldvm_multigpu_check.py > gpurun_out/s3h_ldvm_multigpu_2gpu.txt 2>&1
tail -5 gpurun_out/s3h_ldvm_multigpu_2gpu.txt
