#!/bin/bash
for b in 64 148 185; do
  timeout 100 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | tail -1 > /tmp/l.json
  python - "$b" <<'PY'
import json,sys
d=json.load(open('/tmp/l.json'))
print(sys.argv[1], round(d["value"]/1e6,2), round(d["ms_per_step"],4), round(d["roofline"]["step"]["frac"],4), round(d["roofline"]["fwd_kernel"]["achieved"]), round(d["roofline"]["achieved"]))
PY
done
