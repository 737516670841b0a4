#!/bin/bash
# quick GPU check: parity tests + bench + variant timings.  gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh TAG'
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest.log
timeout 300 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench.json"))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"launches/step",d["gpu_launches_per_step"])
except Exception as e:
    print("bench parse failed",e); print(open("$O/${TAG}_bench.err").read()[-2000:])
PY
timeout 300 python scripts/variant_bench.py C3 4,5 > $O/${TAG}_variants.log 2>&1; cat $O/${TAG}_variants.log
timeout 200 python scripts/stage_times.py C3 > $O/${TAG}_stages.log 2>&1; cat $O/${TAG}_stages.log
