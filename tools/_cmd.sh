set -u
OUT=gpurun_out; TAG=r2r
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_consensus_dropin.py tests/test_gpu_multidev.py -x -q) > $OUT/pytest_online_$TAG.log 2>&1; tail -n 5 $OUT/pytest_online_$TAG.log
(time timeout 600 python bench.py) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -3 $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:.1f} packed {d.get('e2e_packed',{}).get('value')}")
print("online", json.dumps(d.get("online_query"), indent=1))
PY
