for cfg in ${CFGS:-4 7 5 3}; do for lib in "" scripts/microbench/libttneval_other.so; do
  if [ -n "$lib" ]; then export LIBTTNEVAL=$PWD/$lib; else unset LIBTTNEVAL; fi
  python bench.py --config $cfg --steps 6 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('cfg$cfg [$lib]', 'value %.4g kernel_ms %.4f frac %.3f' % (l['value'], l['kernel_ms_events'], l['roofline']['frac']))"
done; done
