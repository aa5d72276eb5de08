import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, smolscale_b200 as sb
wi,hi,wo,ho = 3840,2160,3839,2159
ti,to = 0,0
d_in = torch.randint(0,256,(hi*wi*4,),dtype=torch.uint8,device='cuda')
d_out = torch.zeros(ho*wo*4,dtype=torch.uint8,device='cuda')
for _ in range(3):
    sb.scale_simple(d_in.data_ptr(), ti, wi, hi, wi*4, d_out.data_ptr(), to, wo, ho, wo*4, 0)
torch.cuda.synchronize()
