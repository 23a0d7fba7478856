#!/bin/bash
# A/B of compile-time variants: variants/lib_<name>.so are copied over the library one at a time (the box is scratch)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
cp jaxpm_b200/libjaxpm_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp variants/lib_$v.so jaxpm_b200/libjaxpm_b200.so
  timeout 300 python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --force-mode potential > $OUT/bench_$v.json 2> $OUT/bench_$v.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$v.json"))
    print("$v: ms/step", round(d["ms_per_step"], 4), {k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("$v: failed", e, open("$OUT/bench_$v.err").read()[-500:])
PY
done
cp /tmp/lib_orig.so jaxpm_b200/libjaxpm_b200.so
