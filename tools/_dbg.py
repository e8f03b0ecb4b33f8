import sys; sys.path.insert(0,'.')
import seriation_b200 as S, numpy as np
ds = S.Dataset.synthetic(1024, 4096, 16)
for trial in range(2):
    run = S.Run(ds, 1184, seed=1+trial, store=S.STORE_PI, max_samples=4)
    run.init().sync()
    print('after init check', run.check())
    for call in range(3):
        run.advance(1, True).sync()
        bad = run.check()
        fl = [(i, run.flags(i)) for i in range(1184) if run.flags(i)] if bad else []
        print('trial', trial, 'call', call, 'bad', bad, fl[:8])
        if bad:
            i = fl[0][0]; st = run.state(i)
            print(' a>b?', int((st['a'] > st['b']).sum()), 'amin', st['a'].min(), 'bmax', st['b'].max(), 'tot', st['tot'], 'sumtot', st['tot'].sum(), 1024*4096,
                  'recount', st['t0'].sum(), st['f0'].sum(), st['t1'].sum(), st['f1'].sum(), 'll', st['loglik'])
            break
    run.close()
