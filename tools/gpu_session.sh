#!/usr/bin/env bash
# One gpurun call that gathers everything a round needs to decide what to work on.  Every section
# runs under its own timeout and writes into gpurun_out/ (merged back by gpurun); a section that fails
# does not stop the others.  Copy what should be judged into profiles/ afterwards.
#
#   gpurun --timeout 900 -- 'bash tools/gpu_session.sh all'
#   gpurun --timeout 300 -- 'bash tools/gpu_session.sh tests bench'
#
# sections: tests   pytest -m gpu (the parity tests proper)
#           bench   python bench.py (N=1 line) and the reference arm
#           launch  ncu launch list of a short bench run (per-launch device times: compare SHARES)
#           ncu     ncu --set full of one launch of each hot kernel, raw + source pages exported as CSV
#           sweep   BASELINE configs at full size with their parity checks (tests/config_sweep.py)
#           ab      small-k lookup A/B (tools/mid_tier_ab.py) and the FASTQ-ingest kernels (tools/ingest_sweep.py)
#           peer    the peer-memory multi-rank path with 4 and 8 ranks sharing one device
#           micro   tools/micro: the integer roof (int_roof -> copy to profiles/int_roof.json, bench.py reads it)
#                   and the random-slot atomics benchmark
# No number printed by a run under ncu is a bench value.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
TAG=${TAG:-$(date +%H%M)}
ARGS=" $* "
want() { [[ "$ARGS" == *" all "* ]] || [[ "$ARGS" == *" $SECTION "* ]]; }

SECTION=tests
if want; then
    (time timeout 400 python -m pytest tests -m gpu -x -q --durations=10) > "$OUT/pytest_gpu_$TAG.log" 2>&1
    tail -n 3 "$OUT/pytest_gpu_$TAG.log"
    # experimental kernel variants (default off): parity before any of them is switched on
    (time NSMH_TEST_EXPERIMENTS=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -k experiment) \
        > "$OUT/pytest_experiments_$TAG.log" 2>&1
    tail -n 3 "$OUT/pytest_experiments_$TAG.log"
fi

SECTION=bench
if want; then
    timeout 300 python bench.py > "$OUT/bench_$TAG.json" 2> "$OUT/bench_$TAG.err"
    timeout 200 python bench.py --impl reference > "$OUT/bench_reference_$TAG.json" 2>> "$OUT/bench_$TAG.err"
    cut -c1-400 "$OUT/bench_$TAG.json"
fi

SECTION=launch
if want; then
    timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
        --log-file "$OUT/launches_$TAG.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
        > "$OUT/bench_under_ncu_$TAG.log" 2>&1
    python tools/launch_list.py "$OUT/launches_$TAG.csv" 40 > "$OUT/launches_$TAG.txt" 2>&1
fi

SECTION=ncu
if want; then
    # -s skips the warm-up launches of each kernel, -c 1 captures one launch; ~40 replays each
    for K in sketch_filter_kernel table_insert_kernel count_kernel pack_ascii_kernel sketch_fixup_kernel; do
        timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 3 -c 1 \
            -o "$OUT/prof_${K}_$TAG" -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
            > "$OUT/ncu_${K}_$TAG.log" 2>&1
        if [[ -f "$OUT/prof_${K}_$TAG.ncu-rep" ]]; then
            ncu -i "$OUT/prof_${K}_$TAG.ncu-rep" --page raw --csv > "$OUT/prof_${K}_${TAG}_raw.csv" 2>/dev/null
            ncu -i "$OUT/prof_${K}_$TAG.ncu-rep" --page source --csv > "$OUT/prof_${K}_${TAG}_source.csv" 2>/dev/null
            python tools/ncu_summary.py "$OUT/prof_${K}_${TAG}_raw.csv" > "$OUT/prof_${K}_${TAG}_summary.txt" 2>&1
        fi
    done
fi

SECTION=sweep
if want; then
    timeout 600 python tests/config_sweep.py --out "$OUT/config_sweep_$TAG.jsonl" > "$OUT/config_sweep_$TAG.log" 2>&1
    tail -n 3 "$OUT/config_sweep_$TAG.log"
fi

SECTION=ab
if want; then
    # filter kernel: default phase 2 vs the balanced one (NSMH_SKETCH_BALANCED), device-resident step only
    # ... the lookup with / without the speculative placement (NSMH_LOOKUP_SPECULATE), the fix-up scan with
    # 4 / 8 words per lane and step (NSMH_FIXUP_WIDTH), the filter density (NSMH_LAMBDA_LOG2: a cheaper
    # phase 2 moves the optimum towards fewer fix-ups); last line: everything together
    for V in "0 0 4 2" "1 0 4 2" "0 1 4 2" "0 0 8 2" "0 0 4 3" "1 0 4 3" "1 1 4 3"; do
        set -- $V
        NSMH_SKETCH_BALANCED=$1 NSMH_LOOKUP_SPECULATE=$2 NSMH_FIXUP_WIDTH=$3 NSMH_LAMBDA_LOG2=$4 timeout 120 python bench.py \
            --steps 10 --no-cpu-baseline --no-e2e --no-ingest > "$OUT/bench_bal$1_spec$2_fix$3_lam$4_$TAG.json" \
            2> "$OUT/bench_bal$1_spec$2_fix$3_lam$4_$TAG.err"
        python - "$OUT/bench_bal$1_spec$2_fix$3_lam$4_$TAG.json" "$1" "$2" "$3" "$4" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
p = d["phases_last_step"]
print(f"balanced={sys.argv[2]} speculate={sys.argv[3]} fixup_width={sys.argv[4]} lambda_log2={sys.argv[5]}: "
      f"ms/step {d['ms_per_step']:.4f}  sketch {p['sketch_ms']:.4f} (main kernel {p['sketch_main_kernel_ms']:.4f})  "
      f"query {p['query_ms']:.4f}")
PY
    done
    timeout 120 python tools/mid_tier_ab.py > "$OUT/mid_tier_ab_$TAG.jsonl" 2> "$OUT/mid_tier_ab_$TAG.err"
    timeout 120 python tools/ingest_sweep.py --iters 16 32 64 > "$OUT/ingest_sweep_$TAG.jsonl" 2> "$OUT/ingest_sweep_$TAG.err"
    cat "$OUT/mid_tier_ab_$TAG.jsonl" "$OUT/ingest_sweep_$TAG.jsonl"
fi

SECTION=micro
if want; then
    timeout 120 tools/micro/int_roof > "$OUT/int_roof_$TAG.json" 2> "$OUT/int_roof_$TAG.err"
    cat "$OUT/int_roof_$TAG.json"
    timeout 120 tools/micro/atom_bench > "$OUT/atom_bench_$TAG.txt" 2>&1
fi

SECTION=peer
if want; then
    export NSMH_MG_TIMEOUT_MS=60000
    for W in 4 8; do
        (time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 \
            --master-port 2960$W tools/peer_same_device.py) > "$OUT/peer_same_device_w${W}_$TAG.log" 2>&1
        grep -h "peer-memory" "$OUT/peer_same_device_w${W}_$TAG.log"
    done
fi
exit 0
