CMD="python bench.py --steps 40 --warmup 5 --no-train --no-e2e --no-cpu-baseline --configs none --no-sparse --no-measure-copy"
for W in 8 4 2; do for E in $(seq 0 $((W-1))); do $CMD --emulate-world $W --emulate-rank $E > gpurun_out/r2_emu_${W}_${E}.json 2>/dev/null; done; done
