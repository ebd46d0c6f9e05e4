#!/bin/bash
# One-GPU measurement pass for profiles/ (CUDA-event timed, nothing under a profiler): bash tools/final_measure.sh
mkdir -p gpurun_out/final
o=gpurun_out/final
python bench.py --steps 10 --warmup 3 > $o/r2_bench_c3.json 2> $o/r2_bench_c3.err
python bench.py --impl reference --steps 2 --warmup 1 > $o/r2_bench_c3_reference.json 2> $o/r2_bench_c3_reference.err
python bench.py --aux logits --steps 10 --warmup 3 --no-cpu-baseline > $o/r2_bench_c3_logits_only.json 2> $o/logits.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $o/r2_bench_c4.json 2> $o/c4.err
python bench.py --workload resnet --steps 5 --warmup 3 --no-cpu-baseline > $o/r2_bench_resnet.json 2> $o/resnet.err
python bench.py --hybrid --steps 5 --warmup 3 --no-cpu-baseline > $o/r2_bench_hybrid.json 2> $o/hybrid.err
python bench.py --workload c5 --objective full --steps 10 --warmup 3 --no-cpu-baseline > $o/r2_bench_c5_full.json 2> $o/c5.err
python bench.py --workload c5 --unfrozen --objective full --steps 10 --warmup 3 --no-cpu-baseline > $o/r2_bench_c5_unfrozen.json 2> $o/c5u.err
python bench.py --workload c1 --steps 10 --warmup 3 > $o/r2_bench_c1.json 2> $o/c1.err
python tools/kbench.py > $o/r2_kbench.txt 2>&1
python tools/train_kbench.py > $o/r2_train_kbench.txt 2>&1
python tools/norm_bench.py --sweep > $o/r2_norm_sweep.txt 2>&1
python tools/attn_bench.py > $o/r2_attn_bench.txt 2>&1
python tools/attn_bench.py 64 256 4 128 >> $o/r2_attn_bench.txt 2>&1
for f in $o/r2_bench_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
    print(sys.argv[1].split("/")[-1], round(d.get("value",0),1), d.get("unit"), round(d.get("ms_per_step",0),2), "e2e", round((d.get("e2e") or {}).get("value",0),1), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
except Exception as e: print(sys.argv[1], "ERR", e)
P
done
python tools/stem_bench.py > $o/r2_stem_bench.txt 2>&1
python tools/norm_bench.py > $o/r2_norm_bench.txt 2>&1
