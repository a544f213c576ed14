"""Experiment: time the weight-gradient kernel with parts switched off (D2B_DCN_DEBUG bits: 1 no gather loads, 2 no MMA,
4 no per-stage tap build)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch, math
sys.path.insert(0, %r)
from detectron2_b200 import ops
n,c,h,w=2,128,100,168
g=torch.Generator().manual_seed(0)
x=torch.randn(n,c,h,w,generator=g).cuda(); off=(torch.randn(n,18,h,w,generator=g)*2).cuda()
wt=(torch.randn(c,c,3,3,generator=g)*0.03).cuda(); go=torch.randn(n,c,h,w,generator=g).cuda()
f=lambda: ops.deform_conv_backward_op(x,off,None,wt,go,[1,1],[1,1],[1,1],1,1,False,False,True,1)
for _ in range(3): f()
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): f()
b.record(); torch.cuda.synchronize()
print("weight-grad only (incl. pre-passes): %%.1f us" %% (a.elapsed_time(b)*100))
''' % ROOT
for dbg in (0, 7):
    env = dict(os.environ, D2B_DCN_DEBUG=str(dbg))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    print("dbg=%d" % dbg, r.stdout.strip(), r.stderr.strip()[-200:])
