#!/bin/bash
# A/B of B200_FWD_OPTS settings: usage gpu_ab.sh "<net> <batch> <prec>" "<opts A>" "<opts B>" ...
cfg=($1); shift
for o in "$@"; do
  B200_FWD_OPTS="$o" timeout 300 python bench.py --net ${cfg[0]} --batch ${cfg[1]} --prec ${cfg[2]} --no-cpu-baseline --no-other-configs --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${cfg[0]} ${cfg[2]} [$o] value %.0f ms %.4f' % (d['value'], d['ms_per_step']))"
done
