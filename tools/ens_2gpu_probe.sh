# are member kernels slower when both GPUs of a box run their shard at the same time?
cd ${GRAFT_REPO_ROOT:-.}
echo "--- one after the other"
for r in 0 1; do LOCAL_RANK=$r ENS_WORLD=2 python tools/ensemble_bench.py 2>&1 | grep kernel; done
echo "--- both at once"
for r in 0 1; do (LOCAL_RANK=$r ENS_WORLD=2 python tools/ensemble_bench.py 2>&1 | grep kernel) & done; wait
echo "--- both at once, old strided shards"
for r in 0 1; do (ENS_DEAL=0 LOCAL_RANK=$r ENS_WORLD=2 python tools/ensemble_bench.py 2>&1 | grep kernel) & done; wait
echo "--- bench.py ensemble only"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --ldvm-steps 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d.get('ensemble_C3'))"
