#!/bin/bash
# One short GPU pass: c2 bench line (value / e2e / kernel ms) + the GPU parity tests.  Usage: tools/quick_gpu.sh <tag>
tag=${1:-q}
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
b = json.load(open("gpurun_out/${tag}_bench.json"))
print("value %.0f  ms/step %.4f  e2e %.0f  kernel ms %.4f" % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["roofline"]["ms_per_launch"]))
PY
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
