#!/bin/bash
for f in mvregfus_b200/lib/var_*.so; do echo -n "$f: "; MVF_LIB=$PWD/$f python tools/rl_case.py | tail -1; done
