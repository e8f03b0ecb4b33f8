import sys, time; sys.path.insert(0,".")
import seriation_b200 as S
ds=S.Dataset.synthetic(1024,4096,16)
run=S.Run(ds,592,seed=1,store=S.STORE_PI,max_samples=2)
run.init().sync()
done=0
for calls in (1,1,3,5,10,20,40):
    run.elapsed_ms(reset=True); run.advance(calls, False).sync(); ms=run.elapsed_ms(reset=True); done+=calls
    print("after %3d calls: %.0f sweeps/s over the last %d calls"%(done, 592*calls*10/(ms*1e-3), calls), run.counters(0)[2:7].tolist())
