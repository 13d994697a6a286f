run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-rep3 --no-cpu-baseline --timeline 2>&1 \
    | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('timeline '): print(line.strip())
    elif line.startswith('{'):
        d = json.loads(line); print('ms_per_step', round(d['ms_per_step'], 3), 'e2e_ms', round(d['e2e']['ms_per_step'], 3))
"
}
run CS_PRIO=-1,0,-1
run CS_PRIO=-3,0,-2
run CS_PRIO=-2,0,-2 CS_ACCUM0_MINB=4
for m in 4 5 6; do echo "G2 MINB=$m"; CS_ACCUM0_MINB=$m python tools/run_msm_once.py 20 1 3 | tail -1; done
for m in 4 5; do echo "G1 MINB=$m"; CS_ACCUM0_MINB=$m python tools/run_msm_once.py 20 0 3 | tail -1; done
