"""Enqueue / sync time of one detect_multi call per strip engine, repeated (found a first-use allocation stall of the
piece slots: every slot is now sized on first use). Usage: python tools/mspp_probe.py jakobs1"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jagua_rs_b200 as jg
inst = sys.argv[1]
strips=[]; k=0
while os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", f"{inst}_ms{k}.npz")):
    strips.append(jg.load_workload(f"{inst}_ms{k}")); k+=1
S=len(strips); n_items=len(strips[0].items)
cdes=[jg.build_engine(wl, device=0) for wl in strips]
for c in cdes: c.sync_layout()
per=262144
sizes=jg.split_batch(per, n_items); offs=np.concatenate([[0],np.cumsum(sizes)]).astype(np.int64); ids=list(range(n_items))
ht=[jg.PinnedBuffer((per,4),np.float32) for _ in range(S)]
for s in range(S):
    for i,n in enumerate(sizes): ht[s].array[offs[i]:offs[i+1]]=jg.make_queries(strips[s], i, n, seed=s)
ho=[jg.PinnedBuffer((per,),np.uint8) for _ in range(S)]
for rep in range(6):
    t0=time.perf_counter()
    for s in range(S):
        cdes[s].detect_multi(ids, offs, ht[s].array, out=ho[s].array, wait=False)
    t1=time.perf_counter()
    for c in cdes: c.sync()
    t2=time.perf_counter()
    print(f"rep {rep}: enqueue {1e3*(t1-t0):.3f} ms, sync {1e3*(t2-t1):.3f} ms", [int((ho[s].array==0).sum()) for s in range(S)])
