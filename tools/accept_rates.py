"""acceptance counts per sweep of the pi proposals on the synthetic 1024 x 4096 matrix: python tools/accept_rates.py [chains] [calls per stage] [stages]"""
import sys
sys.path.insert(0, '.')
import numpy as np
import seriation_b200 as S
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 10
stages = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ds = S.Dataset.synthetic(1024, 4096, 16)
run = S.Run(ds, chains, seed=1)
run.init()
prev = np.zeros(8)
for st in range(stages):
    run.advance(calls, False).sync()
    cur = np.mean([run.counters(i) for i in range(0, chains, max(1, chains // 16))], axis=0)
    d = cur - prev
    prev = cur
    print("sweeps %5d..%5d: per sweep  a/b changed %.1f  pi1 %.3f  pi2 %.3f  swap %.3f  pi3 %.3f" %
          (st * calls * 10, (st + 1) * calls * 10, d[2] / d[7], d[3] / d[7], d[4] / d[7], d[5] / d[7], d[6] / d[7]))
