"""One teacher-forced forward of the headline shape (tiny, B=256, N=128, T=768) -- the target of ncu captures."""
import sys
import torch
import efficientspeech_b200 as es
from efficientspeech_b200.params import init_state_dict
from efficientspeech_b200.synthetic import make_batch

variant = sys.argv[1] if len(sys.argv) > 1 else "tiny"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = es.VARIANTS[variant]
model = es.build_model(variant)
es.load_numpy_state(model, init_state_dict(cfg, seed=0))
model = model.to("cuda:0").eval()
model.return_features = False
batch = make_batch(cfg, 256, 128, seed=1000, ragged=False, fixed_duration=6)
x = {k: torch.from_numpy(v).to("cuda:0") for k, v in batch.items()}
x["max_mel_len"] = 768
with torch.no_grad():
    for _ in range(reps):
        out = model(x, train=True)
torch.cuda.synchronize()
model.check_async_errors()
print("ok", tuple(out["mel"].shape))
