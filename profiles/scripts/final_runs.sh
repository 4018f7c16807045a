#!/bin/bash
# Bench lines kept under profiles/ (run on a GPU box through gpurun; copy gpurun_out/r02_bench_*.json to profiles/).
#   bash profiles/scripts/final_runs.sh 1      single-GPU lines: configs[1] (with the CPU oracle leg), configs[2] bf16 / f32, configs[4]
#   bash profiles/scripts/final_runs.sh N      N-GPU lines (N = 2, 4, 8): configs[1] (= configs[3] at 8) and configs[4]
OUT=gpurun_out
mkdir -p $OUT
N=${1:-1}
last_json() { grep '^{' | tail -1; }
if [ "$N" = "1" ]; then
  python bench.py --steps 30 --warmup 5 2>$OUT/r02_bench_c1_1gpu.err | last_json > $OUT/r02_bench_c1_1gpu.json
  python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | last_json > $OUT/r02_bench_c1_reference_arm.json
  python bench.py --config 2 --steps 15 --warmup 3 2>$OUT/r02_bench_c2_bf16.err | last_json > $OUT/r02_bench_c2_bf16.json
  python bench.py --config 2 --dtype f32 --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | last_json > $OUT/r02_bench_c2_f32.json
  python bench.py --config 4 --steps 15 --warmup 3 2>$OUT/r02_bench_c4_1gpu.err | last_json > $OUT/r02_bench_c4_1gpu.json
else
  for c in 1 4; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$c \
        bench.py --config $c --gpus $N --steps 30 --warmup 5 2>$OUT/r02_bench_c${c}_${N}gpu.err | last_json > $OUT/r02_bench_c${c}_${N}gpu.json
  done
fi
for f in $OUT/r02_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d.get("impl", "ours"), d["n_gpus"], d["dtype"], "%.3f ms" % d["ms_per_step"], "%.2f M points/s" % (d["value"] / 1e6))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
