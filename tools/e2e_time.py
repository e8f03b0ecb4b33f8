import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, seriation_b200 as S
from tools.datasets import load_hex_dataset
X,hard=load_hex_dataset('g2s2')
def T(label,t0): 
    print('%-28s %.1f ms'%(label,(time.perf_counter()-t0)*1e3)); return time.perf_counter()
for rep in range(10):
    t=time.perf_counter(); t00=t
    ds=S.Dataset.from_bits(X,hard); t=T('dataset',t)
    run=S.Run(ds,16384,mode=S.MODE_FREE,seed=1,store=S.STORE_PI,max_samples=5); t=T('run create',t)
    run.init().advance(5,False).advance(5,True); print('   gpu ms %.1f'%run.elapsed_ms()); run.sync(); t=T('init+advance+sync',t)
    st=run.chain_stats(); t=T('chain_stats',t)
    ch,_,_=S.select_chains(st['e_negloglik'],4); t=T('select',t)
    cnt=run.po_counts(np.pad(ch,(0,4-len(ch)),constant_values=-1)); t=T('po_counts',t)
    po=S.po_finalize(cnt,4); t=T('po_finalize',t)
    run.close(); t=T('close',t)
    print('total %.1f ms'%((time.perf_counter()-t00)*1e3))
