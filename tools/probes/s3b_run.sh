set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3b_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3b_gputests.log
tail -3 gpurun_out/s3b_gputests.log
for m in 0 1 2 3 4; do UNSFLOW_SWEEP_M=$m python tools/sweep_probe.py >> gpurun_out/s3b_sweep_probe.txt 2>&1; done
python tools/c1_run.py > gpurun_out/s3b_c1.txt 2>&1
python tools/c1_run.py >> gpurun_out/s3b_c1.txt 2>&1
UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 100000 8 > gpurun_out/s3b_stepper_1e5.txt 2>&1
UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 10000 8 >> gpurun_out/s3b_stepper_1e5.txt 2>&1
UNSFLOW_NO_GRAPH=1 C1_STEPS=300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s3b_c1_launches.csv python tools/c1_run.py > gpurun_out/s3b_c1_ncu.log 2>&1
tail -2 gpurun_out/s3b_c1.txt
