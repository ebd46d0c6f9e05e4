run() {
  tag=$1
  python tools/kbench.py --only "gelu" > gpurun_out/ab_kbench_$tag.txt 2>&1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c3_$tag.json 2>/dev/null
  python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c4_$tag.json 2>/dev/null
}
run erf
cp lib/libb200fusion.so /tmp/keep.so; cp lib/libb200fusion_tanh.so lib/libb200fusion.so
run tanh
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_trained_gpu.py -m gpu -q --tb=line 2>&1 | tail -8
cat gpurun_out/trained_parity_report.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:(v['max_rel'],v['argmax_agreement']) for k,v in d.items() if isinstance(v,dict) and 'max_rel' in v})"
cp /tmp/keep.so lib/libb200fusion.so
paste gpurun_out/ab_kbench_erf.txt gpurun_out/ab_kbench_tanh.txt | cut -c1-64,106-140
python - <<'P'
import json
for w in ("c3","c4"):
    for t in ("erf","tanh"):
        d=json.loads([l for l in open(f"gpurun_out/ab_{w}_{t}.json") if l.startswith("{")][0]); print(w,t,round(d["value"]),round(d["ms_per_step"],2),d["clocks"]["sm_mhz"])
P
