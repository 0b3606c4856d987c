import numpy as np, sys
rows=[l.strip().split(',') for l in open(sys.argv[1])]
kind=np.array([int(r[1]) for r in rows]); fl=np.array([float(r[2]) for r in rows]); tiles=np.array([int(r[3]) for r in rows]); ms=np.array([float(r[4]) for r in rows])
names={0:'gemm',1:'diag',2:'tri_f',3:'tri_b',4:'gemv_f',5:'gemv_b',6:'transpose',7:'gather',8:'zero',100:'assemble',101:'other'}
for k in sorted(set(kind)):
    m=kind==k; print(f"{names.get(k,k):10s} n={m.sum():5d} ms={ms[m].sum():9.2f} flops={fl[m].sum():.3e}")
g=kind==0; idx=np.arange(len(rows))
tri_idx=idx[(kind>=2)&(kind<=5)]; f_end=tri_idx.min(); s_beg=tri_idx.max()
for name,(a,b) in {'factor':(0,f_end),'selinv':(s_beg+1,len(rows))}.items():
    m=(idx>=a)&(idx<b)&g
    print(name, "gemm ms %.1f useful flops %.3e TF/s %.1f n %d | diag ms %.1f | all ms %.1f"%(ms[m].sum(), fl[m].sum(), fl[m].sum()/ms[m].sum()/1e9, m.sum(), ms[(idx>=a)&(idx<b)&(kind==1)].sum(), ms[(idx>=a)&(idx<b)].sum()))
    lost=ms[m]-fl[m]/33e9
    for i in idx[m][np.argsort(-lost)[:10]]: print("   idx",i,"tiles",tiles[i],"ms",round(ms[i],3),"TF",round(fl[i]/ms[i]/1e9,1),"Mflop/tile",round(fl[i]/max(tiles[i],1)/1e6,1))

tag=np.array([int(r[5]) if len(r)>5 else 0 for r in rows]); lvl=np.array([int(r[6]) if len(r)>6 else -1 for r in rows])
tn={0:'none',1:'left_update',2:'panel',3:'right_update',4:'schur',5:'trtri_a',6:'trtri_b',7:'yt',8:'z21',9:'z11_ww',10:'z11_yz'}
print("GEMM by step:")
for t in sorted(set(tag[g])):
    m=g&(tag==t); print(f"  {tn.get(t,t):13s} n={m.sum():4d} ms={ms[m].sum():8.1f} useful TF/s={fl[m].sum()/ms[m].sum()/1e9:6.1f} tiles={tiles[m].sum()}")
print("by level (all kinds): level ms gemm_ms diag_ms")
for l in sorted(set(lvl)):
    m=lvl==l; print(f"  L{l:3d} {ms[m].sum():8.1f} {ms[m&g].sum():8.1f} {ms[m&(kind==1)].sum():8.1f}")
