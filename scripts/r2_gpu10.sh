#!/bin/bash
# N1 bench lines: frozen-base LoRA (rank 16), shared-weight forward vs the dense per-coalition merge vs the plain dense model
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/r2_lora_$tag.json 2> gpurun_out/r2_lora_$tag.err; echo "$tag rc=$?"; tail -2 gpurun_out/r2_lora_$tag.err; python - <<PY
import json
for l in open("gpurun_out/r2_lora_$tag.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$tag", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3) if d["e2e"] else None, {k: round(v, 1) for k, v in d["breakdown"].items() if v}, d["clocks"]["sm_mhz"], "K1", d["roofline_aggregate"]["frac"])
PY
}
run shared --lora-rank 16
run dense --lora-rank 16 --lora-path dense
run plain --no-parity --no-throughput-mode
run shared2 --lora-rank 16
