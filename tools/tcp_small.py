import sys, torch
sys.path.insert(0,'/root/repo')
import fgvc_b200
from oracle import oracle as O
g=torch.Generator().manual_seed(0)
H,W,C,T,L=24,40,128,3,6
f=torch.randn(T+1,C,H,W,generator=g).relu()
q,k=f[T][None],f[:T].permute(1,0,2,3)[None].contiguous()
v=torch.rand(1,L,T,H,W,generator=g)
want=O.propagate_port(q,k,v,radius=6,temperature=0.07,topk=10)
got=fgvc_b200.masked_attention_efficient_v2(q.cuda(),k.cuda(),v.cuda(),6,temperature=0.07,topk=10,engine_id=fgvc_b200.ENGINE_PREFILTER,split="f16")
torch.cuda.synchronize()
print("max err",float((got.cpu()-want).abs().max()))
