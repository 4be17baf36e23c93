# usage: bash tests/tools/ab_env.sh VAR v1 v2 ... : stage probe per value of an A/B environment switch.
# The value "-" runs with VAR unset (several switches only test whether the variable exists, so VAR="" still counts as set).
var=$1; shift
for v in "$@"; do
  if [ "$v" = "-" ]; then cmd="env -u $var"; else cmd="env $var=$v"; fi
  $cmd python tests/tools/stage_probe.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['env'], d['ms_per_step'], d['stage_ms'])"
done
