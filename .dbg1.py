import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import lumahdrv_b200 as L
from oracle import pyoracle as po
from test_oracle import adversarial_frame
from conftest import max_ulp
po.build()
for cs in ("LUV","RGB","YCBCR","XYZ"):
    q=L.LumaQuantizer(); q.setQuantizer("PQ",11,cs,8)
    o=po.Oracle().setQuantizer("PQ",11,cs,8)
    for sc in (1.0,0.25):
        f=adversarial_frame(64,16,lut=o.getMapping())
        a,b=f.copy(),f.copy()
        q.transformColorSpace(a,True,sc); o.transformColorSpace(b,True,sc)
        fwd_in=b.copy()
        a=b.copy()
        q.transformColorSpace(a,False,sc); o.transformColorSpace(b,False,sc)
        ai=a.view(np.uint32); bi=b.view(np.uint32)
        nan_both=np.isnan(a)&np.isnan(b)
        bad=np.argwhere((ai!=bi)&~nan_both)
        print(cs,sc,"bad",len(bad),"max_ulp",max_ulp(a,b))
        for c,y,x in bad[:8]:
            print("  ch",c,"in",fwd_in[:,y,x].tolist(),[hex(v) for v in fwd_in[:,y,x].view(np.uint32)],"gpu",a[c,y,x],hex(ai[c,y,x]),"ref",b[c,y,x],hex(bi[c,y,x]))
