"""quick throughput probe: python tools/quick_tput.py <dataset|synthetic> <chains> <calls> [manycd]"""
import sys, time
sys.path.insert(0, '.')
import seriation_b200 as S
from tools.datasets import load_hex_dataset
name, chains, calls = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ds = S.Dataset.synthetic(1024, 4096, 16) if name == 'synthetic' else S.Dataset.from_bits(*load_hex_dataset(name))
run = S.Run(ds, chains, seed=1, store=S.STORE_PI, max_samples=calls, manycd=len(sys.argv) > 4)
run.init().advance(1, False).sync(); run.elapsed_ms(reset=True)
for rep in range(2):
    run.advance(calls, True); ms = run.elapsed_ms(reset=True)
    print(name, 'chains', chains, 'sweeps/chain', calls * 10, 'ms %.1f' % ms, 'sweeps/s %.0f' % (chains * calls * 10 / (ms * 1e-3)))
print('check', run.check())
