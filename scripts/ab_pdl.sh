B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-roofline"
run() { echo -n "$1: "; env $1 $B 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])"; }
run "EB200_NO_PDL=1 EB200_CONV3_NO_2CTA=1"
run "EB200_NO_PDL=1"
run "EB200_PDL_MASK=1 EB200_CONV3_NO_2CTA=1"
run "EB200_PDL_MASK=2 EB200_CONV3_NO_2CTA=1"
run "EB200_PDL_MASK=4 EB200_CONV3_NO_2CTA=1"
run "EB200_PDL_MASK=8 EB200_CONV3_NO_2CTA=1"
run "EB200_PDL_MASK=7 EB200_CONV3_NO_2CTA=1"
