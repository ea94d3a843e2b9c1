# pair-tile share of every rank of an 8-way split, timed on ONE GPU (UNSFLOW_SYM_SHARE: timing probe, partial results)
for R in 8 4; do
  for r in 0 1 2 3 4 5 6 7; do
    echo "== R=$R share $r/8"
    UNSFLOW_SYM_R=$R UNSFLOW_SYM_SHARE=$r,8 UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 100000 4 2>&1 | grep -o "pair tiles [0-9.]* us"
  done
done
echo "== full (1 rank)"
UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 100000 4 2>&1 | grep -o "pair tiles [0-9.]* us"
UNSFLOW_SYM_R=4 UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 100000 4 2>&1 | grep -o "pair tiles [0-9.]* us"
