import sys, os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from rsl_rl_b200 import _native as nat
from rsl_rl_b200.networks.mlp import linear_forward
torch.manual_seed(0)
M,K,N=4096,512,256
x=torch.rand(M,K,device="cuda")+0.5; w=(torch.rand(N,K,device="cuda")+0.5)/K
y=torch.empty(M,N,device="cuda")
linear_forward([(x,w,None,y,0)], nat.PREC_TF32)
ref=(x.double()@w.double().T)
rel=((y.double()-ref)/ref)
print("dtype_sel", os.environ.get("RSLRL_TMA_DTYPE"), "mean rel err (bias)", float(rel.mean()), "max abs rel", float(rel.abs().max()))
# emulate truncation and RN
def trunc(t): return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
def rn(t):
    i=t.view(torch.int32); r=((i + 0x0FFF + ((i>>13)&1)) & ~0x1FFF); return r.view(torch.float32)
for name,f in (("trunc",trunc),("rn",rn)):
    e=(f(x).double()@f(w).double().T); print(name, "mean rel", float(((e-ref)/ref).mean()), " kernel-vs-this max", float(((y.double()-e)/ref).abs().max()))
