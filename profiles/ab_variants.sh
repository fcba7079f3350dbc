for v in _base "" _base ""; do
  PSVO_LIB_PATH=$PWD/psvo_b200/lib/libpsvo_b200$v.so timeout 100 python bench.py --steps 3 --warmup 3 --T 200 --no-e2e --no-cpu-baseline --no-ffbsi 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['clocks'])"
done
