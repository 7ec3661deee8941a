#!/bin/bash
# usage: gpu_sweep.sh TAG  -- runs the env-variable sweeps listed in tools/sweep_TAG.txt ("ENV=.. ENV=.. -- bench args")
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
TAG=$1
: > $O/sweep_$TAG.txt
while IFS= read -r line; do
  [ -z "$line" ] && continue
  envs="${line%%--*}"; args="${line#*--}"
  echo "### $line" >> $O/sweep_$TAG.txt
  env $envs python bench.py --no-cpu --no-extra --no-e2e --steps 1 --warmup 1 $args 2>> $O/sweep_$TAG.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    print('value %.1f ms/step %.1f batch %s' % (d['value'], d['ms_per_step'], d['batch']), {k: round(v,1) for k,v in d['device_ms'].items()}, 'pipe %.3f' % d['roofline']['binding_unit']['frac'])
" >> $O/sweep_$TAG.txt
done < tools/sweep_$TAG.txt
cat $O/sweep_$TAG.txt
