python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/s3j_bench_8gpu.json 2> gpurun_out/s3j_bench_8gpu.err; echo "bench rc=$?"
UNSFLOW_STEP_TRACE=1 MG_WAKE=100000 MG_STEPS=10 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/ldvm_multigpu_check.py > gpurun_out/s3j_ldvm_multigpu_8gpu.txt 2>&1
tail -3 gpurun_out/s3j_ldvm_multigpu_8gpu.txt
