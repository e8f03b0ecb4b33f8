"""where the wall time of a long e2e job goes: python tools/e2e_time.py [chains=4096] [calls=200]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, seriation_b200 as S
from tools.datasets import load_hex_dataset
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
X, hard = load_hex_dataset('g2s2')
def T(label, t0):
    print('%-28s %.1f ms' % (label, (time.perf_counter() - t0) * 1e3)); return time.perf_counter()
for rep in range(2):
    t = time.perf_counter(); t00 = t
    ds = S.Dataset.from_bits(X, hard); t = T('dataset', t)
    run = S.Run(ds, chains, seed=1, store=S.STORE_PI, max_samples=1000); t = T('run create (4.3 GB store)', t)
    run.init().sync(); t = T('init', t)
    run.advance(calls, False).sync(); t = T('burn-in %d calls' % calls, t)
    c = [run.counters(i) for i in range(0, chains, 64)]; t = T('64 x counters()', t)
    run.advance(calls, True).sync(); t = T('sampling %d calls' % calls, t)
    res = run.cross_chain(4); t = T('cross_chain', t)
    ok = run.check(); t = T('check', t)
    run.close(); t = T('close', t)
    print('total %.1f ms' % ((time.perf_counter() - t00) * 1e3))
