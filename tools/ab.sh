#!/bin/bash
# A/B timing of experimental library builds on one GPU box: tools/ab.sh <tag> lib1.so lib2.so ...
# (each is loaded through VOLT_B200_LIB; prints the c2 kernel time and evals/s of every variant)
tag=$1; shift
for lib in "$@"; do
  VOLT_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_$(basename $lib .so).json 2> gpurun_out/${tag}_$(basename $lib .so).err
  python - <<PY
import json
try:
    b = json.load(open("gpurun_out/${tag}_$(basename $lib .so).json"))
    print("$lib: value %.0f  e2e %.0f  kernel ms %.4f" % (b["value"], b["e2e"]["value"], b["roofline"]["ms_per_launch"]))
except Exception as e:
    print("$lib: FAILED", e)
PY
done
