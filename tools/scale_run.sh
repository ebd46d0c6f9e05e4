#!/bin/bash
# Multi-GPU bench lines for profiles/: bash tools/scale_run.sh N ["c3 c5u c5 c4"]  (one box, N ranks over NCCL, 127.0.0.1 rendezvous)
N=$1
mkdir -p gpurun_out
run() {  # name, bench args...
  local name=$1; shift
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 "$@" > gpurun_out/r2_${name}_n${N}.json 2> gpurun_out/r2_${name}_n${N}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $N "$@" > gpurun_out/r2_${name}_n${N}.json 2> gpurun_out/r2_${name}_n${N}.err
  fi
  echo "$name n=$N rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/r2_${name}_n${N}.json | head -2
}
WHICH=${2:-"c3 c5u c5 c4"}
for w in $WHICH; do
  case $w in
    c3) run c3 --steps 10 --warmup 3 --no-cpu-baseline ;;
    c5u) run c5u --workload c5 --unfrozen --objective full --steps 10 --warmup 3 --no-cpu-baseline ;;
    c5) run c5 --workload c5 --objective full --steps 10 --warmup 3 --no-cpu-baseline ;;
    c4) run c4 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline ;;
  esac
done
