# A/B of the chi = 1 table kernel (bench.py --config 6) against another build of the library:
#   git worktree add /tmp/wt <commit> && make -C /tmp/wt/itensornumericalanalysis.jl_b200/csrc -j8 libttneval.so
#   cp /tmp/wt/itensornumericalanalysis.jl_b200/csrc/libttneval.so scripts/microbench/libttneval_other.so
# (used in round 2 to find a 17 % regression: 5 KB more STATIC shared memory moved the 192 KB-table instance from the
#  196 KB shared-memory carveout to the 228 KB one, i.e. from 60 KB of L1 to 28 KB)
for lib in "" scripts/microbench/libttneval_other.so; do
  if [ -n "$lib" ]; then [ -f "$lib" ] || continue; export LIBTTNEVAL=$PWD/$lib; else unset LIBTTNEVAL; fi
  python bench.py --config 6 --steps 30 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; l=json.loads(sys.stdin.read()); print('[$lib]', 'kernel_ms %.4f frac %.3f' % (l['kernel_ms_events'], l['roofline']['frac']))"
done
