export STEPPER_TWO_FAMILIES=1
for r in 0 1 2 3 4 5 6 7; do
  echo "== two families, share $r/8"
  UNSFLOW_SYM_SHARE=$r,8 UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 100000 4 2>&1 | grep -o "pair tiles [0-9.]* us"
done
echo "== full"
UNSFLOW_STEP_TRACE=1 python tools/stepper_launches.py 100000 4 2>&1 | grep "step trace"
