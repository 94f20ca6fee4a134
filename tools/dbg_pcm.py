import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from oracle import avdec
from selkies_b200 import _native as N
from selkies_b200.session import Session
from tests import synth
w,h=128,96
f=synth.desktop(w,h,0); f[:48]=synth.noise(w,48,5)
enc=oracle.RefEncoder(w,h,1)
ref=enc.encode_bgra(f,True,qp=0)
with Session(w,h,rc_mode=N.B2V_RC_CQP,crf=0) as s:
    s.submit(f); s.flush(); got=s.take_frames(); gy,guv=s.recon()
g=got[0].data
print(len(g),len(ref)); print(g[28:60].hex()); print(ref[28:60].hex())
ry,ruv=enc.recon()
print('recon equal', np.array_equal(gy,ry), 'rows differing', np.where((gy!=ry).any(axis=1))[0][:10])
y,uv=oracle.csc_nv12(f)
print('gpu recon == source (top)', np.array_equal(gy[:48,:w], y[:48]))
try:
    d=avdec.decode_stream([g]); print('decoded', len(d))
except Exception as e: print('decode error', e)
