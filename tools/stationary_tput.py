"""throughput after burn-in: python tools/stationary_tput.py <dataset> <chains> [burn_calls=300] [calls=20]"""
import sys
sys.path.insert(0, '.')
import seriation_b200 as S
from tools.datasets import load_hex_dataset
name, chains = sys.argv[1], int(sys.argv[2])
burn = int(sys.argv[3]) if len(sys.argv) > 3 else 300
calls = int(sys.argv[4]) if len(sys.argv) > 4 else 20
ds = S.Dataset.from_bits(*load_hex_dataset(name))
run = S.Run(ds, chains, seed=1)
run.init().advance(burn, False).sync()
run.sweep_time(reset=True)
for rep in range(2):
    run.advance(calls, False).sync()
    ms, n = run.sweep_time(reset=True)
    print(name, 'chains', chains, 'after', burn, 'burn-in calls:', '%.0f sweeps/s' % (chains * calls * 10 / (ms * 1e-3)))
