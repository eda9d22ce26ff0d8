#!/bin/bash
for f in mvregfus_b200/lib/var_*.so; do
  echo -n "$f: "
  MVF_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
done
