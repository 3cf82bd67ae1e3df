import torch, sys
sys.path.insert(0,'/root/repo')
from sct_gan_b200 import kernels as kn
B,H,L,d=32,8,1024,768
qkv=torch.randn(B*L,3*d,device='cuda').bfloat16()
o,lse=kn.attn_fwd(qkv[:,:d],qkv[:,d:2*d],qkv[:,2*d:],B,H,L,L,p_drop=0.3,seed=1,offset=1)
do=torch.randn(B*L,d,device='cuda').bfloat16()
dqkv=torch.empty_like(qkv)
for i in range(2):
    kn.attn_bwd(qkv[:,:d],qkv[:,d:2*d],qkv[:,2*d:],o,do,lse,B,H,L,L,dqkv[:,:d],dqkv[:,d:2*d],dqkv[:,2*d:],p_drop=0.3,seed=1,offset=1)
    torch.cuda.synchronize()
    print('---')
