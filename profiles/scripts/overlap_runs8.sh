# 8-GPU A/B of the overlapped gradient all-reduce (DGCNN_OVERLAP_AR=1 default / 0 = single all-reduce in apply_gradient)
cd $GRAFT_REPO_ROOT
for m in 1 0 1 0; do
DGCNN_OVERLAP_AR=$m timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ovl_bench8_$m.json 2> gpurun_out/ovl_bench8_$m.err; echo "bench rc=$?"
python -c "
import json,sys
for l in open('gpurun_out/ovl_bench8_$m.json'):
    if l.startswith('{'):
        d=json.loads(l); print('OVL=$m', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
" | tee -a gpurun_out/ovl_bench8_summary.txt
done
