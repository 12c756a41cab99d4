for lib in "" scripts/microbench/libttneval_other.so "" scripts/microbench/libttneval_other.so; do
  if [ -n "$lib" ]; then export LIBTTNEVAL=$PWD/$lib; else unset LIBTTNEVAL; fi
  python bench.py --steps 10 --no-side-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); e=l['e2e']; print('[$lib]', 'value %.3f e2e %.3f ceiling %.3f pageable %.3f' % (l['value']/1e9, e['value']/1e9, e['copy_ceiling']['points_per_s']/1e9, e['pageable']['value']/1e9))"
done
