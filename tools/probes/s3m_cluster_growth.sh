export GROW_STEPS=3000 GROW_WINDOW=250
echo "== no cluster"; UNSFLOW_NO_CLUSTER=1 python tools/ldvm_bench.py 2>&1 | grep -v "^total"
echo "== cluster 8"; UNSFLOW_CLUSTER_MAX_N=2048 UNSFLOW_CLUSTER_SIZE=8 python tools/ldvm_bench.py 2>&1 | grep -v "^total"
echo "== cluster 16"; UNSFLOW_CLUSTER_MAX_N=2048 UNSFLOW_CLUSTER_SIZE=16 python tools/ldvm_bench.py 2>&1 | grep -v "^total"
