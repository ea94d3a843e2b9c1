cd $GRAFT_REPO_ROOT
export ENS_PYTEST=1 ENS_OUT=r3_ens_queues.txt
ENS_CFGS="0:256:3:1:96 0:256:3:8:96" bash tools/ens_matrix.sh
for r in 3 7; do ENS_PYTEST= ENS_RANK=$r ENS_OUT=r3_ens_queues_rank$r.txt ENS_CFGS="0:256:3:8:96" bash tools/ens_matrix.sh; done
echo "--- global queue (1 queue), old strided shard"
UNSFLOW_ENS_QUEUES=1 ENS_DEAL=0 ENS_PYTEST= ENS_OUT=r3_ens_queues_old.txt ENS_CFGS="0:256:3:8:96 0:256:3:1:96" bash tools/ens_matrix.sh
echo "--- global queue, dealt shard"
UNSFLOW_ENS_QUEUES=1 ENS_PYTEST= ENS_OUT=r3_ens_queues_old2.txt ENS_CFGS="0:256:3:8:96" bash tools/ens_matrix.sh
echo "--- 4 GPUs and 2 GPUs share"
ENS_PYTEST= ENS_OUT=r3_ens_queues_w4.txt ENS_CFGS="0:256:3:4:96 0:256:3:2:96" bash tools/ens_matrix.sh
