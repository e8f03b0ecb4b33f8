import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, seriation_b200 as S
from conftest import load_hex_dataset
g = np.load('tests/golden/ref_free_g10s10.npz')
X, hard = load_hex_dataset('g10s10')
ds = S.Dataset.from_bits(X, hard)
for seed in (12345, 7, 2024):
    batch = S.run_all_chains(ds, 100, 1000, 1000, seed=seed, store=S.STORE_PI)
    e = batch.stats()['e_negloglik']
    chosen = S.choose_chains(batch, 8)
    po = S.compute_pair_order_matrix(batch, chosen, 8, 124)
    d = np.abs(po - g['po'])
    ec, ed = S.compute_exp_cd(batch, chosen, 8)
    print('seed', seed, 'min %.1f sigma %.1f (ref %.1f %.1f)' % (e.min(), e.std(), g['e_negloglik'].min(), g['e_negloglik'].std()),
          'sel mean %.1f (ref %.1f)' % (e[chosen].mean(), g['e_negloglik'][g['chosen']].mean()),
          'PO diff max %.4f mean %.5f p99 %.4f frac>0.02 %.4f' % (d.max(), d.mean(), np.percentile(d, 99), (d > 0.02).mean()),
          'Ec %.5f Ed %.5f corr %.4f' % (ec, ed, S.compute_exp_ages(batch, chosen, 8, 124)))
    batch.run.close()
pos = []
for seed in (1, 2, 3, 4):
    batch = S.run_all_chains(ds, 100, 1000, 1000, seed=seed, store=S.STORE_PI)
    pos.append(S.compute_pair_order_matrix(batch, S.choose_chains(batch, 8), 8, 124)); batch.run.close()
for i in range(4):
    for j in range(i + 1, 4):
        d = np.abs(pos[i] - pos[j]); print('gpu-gpu', i, j, 'max %.4f mean %.5f p99 %.4f frac>0.02 %.4f' % (d.max(), d.mean(), np.percentile(d, 99), (d > 0.02).mean()))
    d = np.abs(pos[i] - g['po']); print('gpu-ref', i, 'max %.4f mean %.5f p99 %.4f frac>0.02 %.4f' % (d.max(), d.mean(), np.percentile(d, 99), (d > 0.02).mean()))
# a much larger ensemble: 4096 chains, best 256 -> the Monte-Carlo error of the PO estimate shrinks
batch = S.run_all_chains(ds, 4096, 1000, 1000, seed=5, store=S.STORE_PI)
ch = S.choose_chains(batch, 256); pa = S.compute_pair_order_matrix(batch, ch, len(ch), 124, faithful=False); batch.run.close()
batch = S.run_all_chains(ds, 4096, 1000, 1000, seed=6, store=S.STORE_PI)
ch = S.choose_chains(batch, 256); pb = S.compute_pair_order_matrix(batch, ch, len(ch), 124, faithful=False); batch.run.close()
d = np.abs(pa - pb); print('gpu 256-of-4096 vs same, other seed: max %.4f mean %.5f' % (d.max(), d.mean()))
d = np.abs(pa - g['po']); print('gpu 256-of-4096 vs ref 8-of-100: max %.4f mean %.5f' % (d.max(), d.mean()))
