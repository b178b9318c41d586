#!/bin/bash
for v in "HSIMAE_PDL=1" "HSIMAE_PDL=0" "HSIMAE_PDL=0 HSIMAE_OVERLAP=0"; do
  echo "== $v"; env $v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
done
