import sys, numpy
sys.path.insert(0,'.')
import wendy_b200
from bench import sech2_ic
for n in (int(2e7),int(1e8)):
    x,v,m=sech2_ic(n,2)
    st=wendy_b200.ApproxState(x,v,m,omega2=1.21)
    for call in range(6):
        st.step(1e-3,10)
        c,s=st.layout()
        print(call, st.stats(), 'sum',c.sum(),'max',c.max(),'argmax',c.argmax(), 'min', c.min(), 'std', c.std(), 'hist>200', (c>200).sum(), flush=True)
        if c.max()>215:
            b=c.argmax(); print('  around', b, c[max(0,b-5):b+6], s[max(0,b-2):b+3])
