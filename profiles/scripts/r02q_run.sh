# final single-GPU pass of round 2: GPU test suite, BN-backward microbench, bench line, launch list, ncu --set full of the
# BN-backward kernels of the head (the kernels that changed last)
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > $O/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02q_pytest.log
timeout 60 python profiles/scripts/bn_bwd_bench.py > $O/r02q_bn_bwd.txt 2>&1; cat $O/r02q_bn_bwd.txt
timeout 200 python bench.py --steps 30 --warmup 5 2>$O/r02q_bench.err | grep '^{' | tail -1 > $O/r02q_bench_c1_1gpu.json; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/r02q_bench_c1_1gpu.json')); print('bench', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
DGCNN_CUDA_GRAPH=0 DGCNN_ASYNC_DW=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file $O/r02q_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r02q_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/scripts/step_breakdown.py $O/r02q_launches.csv 70 > $O/r02q_launches_summary.txt 2>&1; head -12 $O/r02q_launches_summary.txt
timeout 200 ncu --profile-from-start off --set full --clock-control none -k regex:"bn_colsum_vec|bn_act_bwd" -c 14 -o /tmp/r02q_bn python profiles/scripts/prof_step.py > $O/r02q_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/r02q_bn.ncu-rep --page raw --csv 2>/dev/null | python profiles/scripts/ncu_summary.py > $O/r02q_bn_kernels_ncu_full.csv; wc -l $O/r02q_bn_kernels_ncu_full.csv
