# usage: bash scripts/ab_env.sh "ENV_A=1" "ENV_B=1 ..."   -> ms/step of the graph-replayed step, variants alternated twice
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
run() { echo -n "[$1]: "; env $1 $B 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])"; }
for rep in 1 2; do for v in "$@"; do run "$v"; done; done
