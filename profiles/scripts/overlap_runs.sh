# 2-GPU check of the overlapped gradient all-reduce: the worker's parity asserts, then bench.py with and without it
set -x
cd $GRAFT_REPO_ROOT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/overlap_worker.py > gpurun_out/ovl_worker2.log 2>&1; echo "worker2 rc=$?"
grep OVERLAP gpurun_out/ovl_worker2.log
for m in 1 0 1 0; do
DGCNN_OVERLAP_AR=$m timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ovl_bench2_$m.json 2> gpurun_out/ovl_bench2_$m.err; echo "bench rc=$?"
python -c "
import json,sys
for l in open('gpurun_out/ovl_bench2_$m.json'):
    if l.startswith('{'):
        d=json.loads(l); print('OVL=$m', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
"
done
