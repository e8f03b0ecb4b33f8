"""One process driving 1 / 2 / 4 / 8 GPUs through ser_multi_* (peer stores over NVLink): the bench step
(init + 5 burn-in + 5 sampling calls x 10 sweeps + selection k = 4 + pair-order counts) on 16 384 g2s2 chains in total.
    python tools/multi_scaling.py [chains_total] [steps]        (needs a box with several GPUs: gpurun --gpus 8)"""
import sys
import time
sys.path.insert(0, '.')
import seriation_b200 as S
from tools.datasets import load_hex_dataset
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ds = S.Dataset.from_bits(*load_hex_dataset("g2s2"))
for g in (1, 2, 4, 8):
    try:
        m = S.Multi(ds, chains, g, seed=20060206, store=S.STORE_PI, max_samples=5)
    except S.SeriationError as e:  # fewer devices than g
        print("stopping at %d GPUs: %s" % (g, e))
        break
    def step():
        m.init().advance(5, 5)
        return m.cross_chain(4)
    for _ in range(3):
        step()
    m.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        r = step()  # reads the result: host-synchronous
    dt = (time.perf_counter() - t0) / steps
    print("ser_multi g2s2 %d chains, %d GPUs: %.1f ms/step -> %.2f M sweeps/s; peer stores %s, chosen %s"
          % (chains, g, dt * 1e3, chains * 100 / dt / 1e6, m.layout()["peer_stores"], r["chosen"].tolist()))
    m.close()
