# cfg2 device-resident rate for team counts of the team-sorted kernel; usage: v6_probe.sh "3 4" [lib]
for v in $1; do
  LIBTTNEVAL=${2:+$PWD/$2} TTN_MMA_V6=$v python bench.py --steps 10 --no-side-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('V6=$v lib=$2 value %.3f G kernel %.3f ms' % (l['value']/1e9, l['kernel_ms_events']))"
done
