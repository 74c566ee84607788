for c in 1 2 4; do
  timeout 300 python bench.py --no-cpu --no-projection --e2e-chunk $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk', $c, d['e2e']['value'], d['e2e']['ms_per_step'])"
done
