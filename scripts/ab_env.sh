#!/usr/bin/env bash
# same-box A/B of environment switches: scripts/ab_env.sh "VAR=VAL" "VAR2=VAL2 VAR3=VAL3" ...   ("" = baseline)
for cfg in "" "$@"; do
  r=$(env $cfg timeout 120 python bench.py --steps 12 --no-cpu-baseline --no-stock-torch --no-roofline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f ms  %.1f img/s' % (d['ms_per_step'], d['value']))")
  echo "[${cfg:-baseline}] $r"
done
