"""One forward of the HiFi-GAN V2 generator (seeded weights) -- the target of ncu captures.  args: B T"""
import sys
import torch
import efficientspeech_b200 as es

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 192
cfg = {"resblock": "1", "upsample_rates": [8, 8, 2, 2], "upsample_kernel_sizes": [16, 16, 4, 4], "upsample_initial_channel": 128,
       "resblock_kernel_sizes": [3, 7, 11], "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]]}
torch.manual_seed(0)
G = es.hifigan.Generator(es.hifigan.AttrDict(cfg)).eval()
G.remove_weight_norm()
G = G.to("cuda:0")
mel = torch.randn(B, 80, T, device="cuda:0") * 1.5 - 4
with torch.no_grad():
    for _ in range(2):
        wav = G(mel)
torch.cuda.synchronize()
print("ok", tuple(wav.shape), float(wav.abs().max()))
