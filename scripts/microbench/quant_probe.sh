# end-to-end rate of config 2 from PINNED host buffers for the host-side quantisation modes (TTN_HOST_QUANT) and the
# image host-buffer calls run (TTN_MMA_LIGHT); usage: quant_probe.sh "light:quant ..." (x:x = the defaults)
for lm in ${1:-1:1 0:1 0:3 0:4 0:2 1:3}; do
  l=${lm%%:*}; m=${lm##*:}
  if [ "$l" = x ]; then unset TTN_MMA_LIGHT TTN_HOST_QUANT; else export TTN_MMA_LIGHT=$l TTN_HOST_QUANT=$m; fi
  python bench.py --steps 10 --no-side-configs --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); e=l['e2e']; print('TTN_MMA_LIGHT=$l TTN_HOST_QUANT=$m value %.3f e2e %.3f ceiling %.3f h2d B/pt %.1f pageable %.3f' % (l['value']/1e9, e['value']/1e9, e['copy_ceiling']['points_per_s']/1e9, e['h2d_bytes_per_step']/l['config']['points_per_gpu'], e['pageable']['value']/1e9))"
done
