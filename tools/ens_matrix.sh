# usage: ENS_CFGS="sym:threads:occ:world ..." bash tools/ens_matrix.sh   (A/B matrix of the ensemble kernel variants on one GPU)
cd ${GRAFT_REPO_ROOT:-.}
if [ -n "$ENS_PYTEST" ]; then python -m pytest tests/test_gpu_ldvm.py -x -q -k "ensemble" > gpurun_out/r3_ens_pytest.log 2>&1; tail -3 gpurun_out/r3_ens_pytest.log; fi
for cfg in $ENS_CFGS; do
  export ENS_RANK=${ENS_RANK:-0}
  IFS=: read sym thr occ world smin <<< "$cfg"
  echo "== SYM=$sym THREADS=$thr OCC=$occ WORLD=$world SYM_MIN=$smin"
  UNSFLOW_ENS_PROFILE=1 UNSFLOW_ENS_SYM_MIN=${smin:-96} UNSFLOW_ENS_SYM=$sym UNSFLOW_ENS_THREADS=$thr UNSFLOW_ENS_OCC=$occ ENS_WORLD=$world python tools/ensemble_bench.py 2>&1 | grep "profile\|kernel"
done 2>&1 | tee gpurun_out/${ENS_OUT:-r3_ens_matrix.txt}
