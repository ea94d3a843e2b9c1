python -m pytest tests/test_gpu_ldvm.py -m gpu -x -q > gpurun_out/s3l_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3l_gputests.log; tail -15 gpurun_out/s3l_gputests.log
for c in 8 16; do echo "== cluster $c"; UNSFLOW_CLUSTER_SIZE=$c python tools/c1_run.py; UNSFLOW_CLUSTER_SIZE=$c python tools/c1_run.py; done
echo "== no cluster"; UNSFLOW_NO_CLUSTER=1 python tools/c1_run.py
for n in 500 1000 1500 2000; do
  for c in 8 16; do echo "== N $n cluster $c"; UNSFLOW_CLUSTER_MAX_N=2048 UNSFLOW_CLUSTER_SIZE=$c python tools/stepper_launches.py $n 64 2>&1 | tail -1; done
  echo "== N $n no cluster"; UNSFLOW_NO_CLUSTER=1 python tools/stepper_launches.py $n 64 2>&1 | tail -1
done
