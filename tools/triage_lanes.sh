# stability of the default configuration under concurrent replay: the bench several times in a row, each in its own process
cd $GRAFT_REPO_ROOT
for i in 1 2 3 4; do
  timeout 100 python bench.py --no-train --no-cpu-baseline --steps 50 2>&1 | grep -v CUDAEvent | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print('OK', d['value'], d['e2e']['value'], d['value_lanes']['value'])
except Exception as e:
    print('FAIL', [l for l in t.splitlines() if 'Error' in l][:3])
"
done
