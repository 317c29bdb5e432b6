import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from common import *
from xgc_msm_b200 import abi
case=make_case(1500,7,network_mode='surrogate')
ha=load_handle([case],1971); hb=load_handle([case],1971); o=load_oracle(case,1971,0)
for at in range(2,40):
    ha.step(at); o.step(at,K.M_ALL)
    for m in P.MODULE_ORDER: hb.step(at,P.MODULE_BIT[m])
    ca,cb,co=ha.counters(0,at,at)[0],hb.counters(0,at,at)[0],o.counters(at)
    if not np.array_equal(ca,co): print('week',at,'FUSED != oracle', np.flatnonzero(ca!=co)[:10])
    if not np.array_equal(cb,co):
        bad=np.flatnonzero(cb!=co); print('week',at,'MODWISE != oracle',bad[:20], cb[bad][:10], co[bad][:10])
        for nme in abi.names('attr'):
            x,y=hb.download_attr(0,nme),o.get_attr(nme)
            if x.shape!=y.shape or not np.array_equal(x,y): print(' attr',nme, x.shape,y.shape, np.flatnonzero(x!=y)[:5] if x.shape==y.shape else '', x[x!=y][:5] if x.shape==y.shape else '', y[x!=y][:5] if x.shape==y.shape else '')
        break
else: print('identical')
