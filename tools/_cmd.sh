set -u
OUT=gpurun_out; TAG=r2i
(time timeout 600 python bench.py) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -5 $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
p = d["phases_last_step"]
print(f"value {d['value']:.1f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:.1f} ({d['e2e']['ms_per_step']:.2f} ms)")
print("e2e_packed", json.dumps(d.get("e2e_packed")))
print("roofline", json.dumps({k:v for k,v in d["roofline"].items() if k in ("frac","binding_resource","issue_slots","kernel_ms")}))
print("roofline_brute", json.dumps(d.get("roofline_brute")))
for k,v in (d.get("legs") or {}).items():
    print(k, json.dumps(v)[:900])
PY
