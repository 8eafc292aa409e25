set -u
OUT=gpurun_out; TAG=r2g
(timeout 600 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu_$TAG.log 2>&1; tail -n 3 $OUT/pytest_gpu_$TAG.log
timeout 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-ingest > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
p = d["phases_last_step"]
print(f"ms/step {d['ms_per_step']:.4f} pack {p['pack_ms']:.4f} sketch {p['sketch_ms']:.4f} (main {p['sketch_main_kernel_ms']:.4f}) build {p['build_ms']:.4f} query {p['query_ms']:.4f} launches {d['gpu_launches']}")
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ingest > $OUT/bench_under_ncu_$TAG.log 2>&1
python tools/launch_list.py $OUT/launches_$TAG.csv 40 | tail -14
