import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, smolscale_b200 as sb
a = sys.argv[1:]
wi, hi, wo, ho = [int(v) for v in a[0:4]] if len(a) >= 4 else (3840, 2160, 3839, 2159)
ti, to, srgb = [int(v) for v in a[4:7]] if len(a) >= 7 else (0, 0, 0)
bi, bo = (3 if ti >= 8 else 4), (3 if to >= 8 else 4)
d_in = torch.randint(0, 256, (hi * wi * bi,), dtype=torch.uint8, device='cuda')
d_out = torch.zeros(ho * wo * bo, dtype=torch.uint8, device='cuda')
for _ in range(3):
    sb.scale_simple(d_in.data_ptr(), ti, wi, hi, wi * bi, d_out.data_ptr(), to, wo, ho, wo * bo, srgb)
torch.cuda.synchronize()
