#!/bin/bash
# A/B of ya_set_priority (YA_PRIO=<levels>): host golden tests with it on, then the bench's e2e / value with 0, 4 and 2 levels.
O=gpurun_out; mkdir -p $O
YA_PRIO=4 python -m pytest tests/test_host_sam.py tests/test_abi.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for lv in 0 4 2 0 4; do
  YA_PRIO=$lv python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $O/prio_$lv.json 2> $O/prio_$lv.err
  python - <<PY
import json
d=json.loads(open("$O/prio_$lv.json").read().strip().splitlines()[-1])
print("YA_PRIO=$lv e2e", round(d["e2e"]["value"]), "value", round(d["value"]), "value@e2ecfg", round(d["value_at_e2e_config"]["value_this_rank"]), "frac", round(d["roofline"]["frac"],3))
PY
done
