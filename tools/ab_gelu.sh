#!/bin/bash
# A/B of the epilogue GELU on ONE box: the product library (MUFU.TANH form) against the same tree built with the
# polynomial-erf GELU.  Build the second library first (in the authoring container, no GPU needed):
#     B200_EXTRA_NVCC_FLAGS=-DB200_GELU_ERF bash <pkg>/csrc/build.sh --clean && cp lib/libb200fusion.so lib/libb200fusion_erf.so
#     bash <pkg>/csrc/build.sh --clean
# then on the GPU box:  bash tools/ab_gelu.sh      (results in gpurun_out/ab_*; round-2 numbers: DESIGN.md 4.1.1)
set -u
[ -f lib/libb200fusion_erf.so ] || { echo "lib/libb200fusion_erf.so missing (see the header of this script)"; exit 1; }
mkdir -p gpurun_out
run() {
  tag=$1
  python tools/kbench.py --only "gelu" > gpurun_out/ab_kbench_$tag.txt 2>&1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c3_$tag.json 2>/dev/null
  python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c4_$tag.json 2>/dev/null
}
run tanh
cp lib/libb200fusion.so /tmp/keep.so; cp lib/libb200fusion_erf.so lib/libb200fusion.so
run erf
cp /tmp/keep.so lib/libb200fusion.so
paste gpurun_out/ab_kbench_erf.txt gpurun_out/ab_kbench_tanh.txt | cut -c1-64,106-140
python - <<'P'
import json
for w in ("c3", "c4"):
    for t in ("erf", "tanh"):
        d = json.loads([l for l in open(f"gpurun_out/ab_{w}_{t}.json") if l.startswith("{")][0])
        print(w, t, round(d["value"]), round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"])
P
