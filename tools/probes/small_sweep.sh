set -e
for T in 256 128; do
  python -c "
from blokus_rl_b200 import build
build.build(force=True, extra_flags=['-DBLK_SMALL_T=$T'])" 
  for E in 65536 1048576; do
    for M in bytes bits; do
      python bench.py --board 7 --players 2 --envs $E --mask $M --steps 300 --warmup 10 --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('T=$T E=$E $M', '%.3e steps/s'%d['value'], 'ms %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'])"
    done
  done
done
