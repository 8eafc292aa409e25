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
#           ab      the remaining switches (speculative placement, table slots per read, tile size), small-k lookup A/B
#                   (tools/mid_tier_ab.py) and the FASTQ-ingest kernels (tools/ingest_sweep.py)
#           peer    the peer-memory multi-rank path with 4 and 8 ranks sharing one device
#           sanitize compute-sanitizer memcheck / initcheck / racecheck over tools/sanitize_smoke.py (every kernel family)
#           e2e     tools/e2e_breakdown.py: host wall time of every call of the host-buffer path, the pipelined calls, the bare copy
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
    # host flows that are not the default (NSMH_LOOKUP_SPECULATE=0 etc.)
    (time NSMH_TEST_EXPERIMENTS=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -k experiment) \
        > "$OUT/pytest_experiments_$TAG.log" 2>&1
    tail -n 3 "$OUT/pytest_experiments_$TAG.log"
    # the C++ drop-in under OpenMP: per-call latency of the online query next to the reference's
    if [[ -x oracle/_ref/dropin_test ]]; then
        python - <<'PY' > "$OUT/dropin_latency_$TAG.log" 2>&1
import os, struct, subprocess, tempfile
import numpy as np
import nanospring_b200 as ns
lengths = ns.synth_lengths(4000, 6000, seed=21)
rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=400_000, genome_seed=3, read_seed=4, p_ins=0.01, p_del=0.01, p_sub=0.02))
td = tempfile.mkdtemp()
p = os.path.join(td, "reads.bin")
with open(p, "wb") as f:
    f.write(struct.pack("<I", rd.offsets.size - 1)); f.write(rd.offsets.astype(np.uint64).tobytes()); f.write(rd.bases.tobytes())
for thr in (1, 8, 32):
    r = subprocess.run(["oracle/_ref/dropin_test", p, "23", "60", "6", td], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(thr)))
    print(r.stdout.strip()); print(r.stderr.strip()[-300:])
PY
        cat "$OUT/dropin_latency_$TAG.log"
    fi
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
        --log-file "$OUT/launches_$TAG.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-legs \
        > "$OUT/bench_under_ncu_$TAG.log" 2>&1
    python tools/launch_list.py "$OUT/launches_$TAG.csv" 40 > "$OUT/launches_$TAG.txt" 2>&1
fi

SECTION=ncu
if want; then
    # -s skips the warm-up launches of each kernel, -c 1 captures one launch; ~40 replays each
    for K in sketch_filter_kernel table_insert_kernel count_kernel pack_ascii_kernel sketch_fixup_kernel; do
        timeout 200 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 3 -c 1 \
            -o "$OUT/prof_${K}_$TAG" -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-legs \
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
    # switches that still exist: speculative placement of the lookup off / on, table slots per read, sketch tile size
    for V in "1 2 640" "0 2 640" "1 4 640" "1 2 512"; do
        set -- $V
        NSMH_LOOKUP_SPECULATE=$1 NSMH_TABLE_SLOTS_PER_READ=$2 NSMH_TILE_WORDS=$3 timeout 120 python bench.py \
            --steps 10 --no-cpu-baseline --no-e2e --no-ingest --no-legs > "$OUT/bench_spec$1_slots$2_tile$3_$TAG.json" \
            2> "$OUT/bench_spec$1_slots$2_tile$3_$TAG.err"
        python - "$OUT/bench_spec$1_slots$2_tile$3_$TAG.json" "$1" "$2" "$3" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
p = d["phases_last_step"]
print(f"speculate={sys.argv[2]} slots_per_read={sys.argv[3]} tile_words={sys.argv[4]}: ms/step {d['ms_per_step']:.4f}  "
      f"sketch {p['sketch_ms']:.4f} (main kernel {p['sketch_main_kernel_ms']:.4f})  build {p['build_ms']:.4f}  query {p['query_ms']:.4f}")
PY
    done
    # nsmh_sketch_build (fix-up pass beside the table insert) against the separate calls
    for F in "" "--separate-build"; do
        timeout 120 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-ingest --no-legs $F 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); p = d['phases_last_step']
print('$F'.strip() or 'nsmh_sketch_build', 'ms/step %.4f' % d['ms_per_step'], 'sketch %.4f build %.4f query %.4f' % (p['sketch_ms'], p['build_ms'], p['query_ms']))"
    done
    timeout 120 python tools/mid_tier_ab.py > "$OUT/mid_tier_ab_$TAG.jsonl" 2> "$OUT/mid_tier_ab_$TAG.err"
    timeout 120 python tools/ingest_sweep.py --iters 16 32 64 > "$OUT/ingest_sweep_$TAG.jsonl" 2> "$OUT/ingest_sweep_$TAG.err"
    cat "$OUT/mid_tier_ab_$TAG.jsonl" "$OUT/ingest_sweep_$TAG.jsonl"
fi

SECTION=sanitize
if want; then
    for T in memcheck initcheck; do
        timeout 300 compute-sanitizer --tool $T --error-exitcode 1 python tools/sanitize_smoke.py > "$OUT/sanitize_${T}_$TAG.log" 2>&1
        echo "$T rc=$?"; tail -n 2 "$OUT/sanitize_${T}_$TAG.log"
    done
    timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 400 python tools/sanitize_smoke.py \
        > "$OUT/sanitize_racecheck_$TAG.log" 2>&1
    grep "and \(Read\|Write\) access" "$OUT/sanitize_racecheck_$TAG.log" | sed 's/0x[0-9a-f]*//g; s/\[[0-9]* hazards\]//' | sort | uniq -c
fi

SECTION=e2e
if want; then
    timeout 200 python tools/e2e_breakdown.py > "$OUT/e2e_breakdown_$TAG.json" 2> "$OUT/e2e_breakdown_$TAG.err"
    cat "$OUT/e2e_breakdown_$TAG.json"
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
