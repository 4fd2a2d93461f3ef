import sys; sys.path.insert(0, '.')
import numpy as np, torch
from oracle import unet_ref as R
from unet_b200.engine import UNetEngine
shape=(64,64,3)
specs = R.layer_specs(shape,1,0.2,True); P = R.init_params(specs, seed=3, trained_like=True)
x,y = R.synthetic_batch(4,64,64,3,1,seed=10)
xd,yd = torch.tensor(x,device='cuda'), torch.tensor(y,device='cuda')
res={}
for dt in ("fp32","bf16"):
    e = UNetEngine(shape, dropout_rate=0.2, dtype=dt); e.dropout_masks_from_step=False; e.set_weights(P)
    o = e.train_forward_backward(xd,yd).cpu().numpy()
    res[dt] = {n: e.wview(n,e.g).cpu().numpy().astype(np.float64) for n,p in e.spec.params.items() if p.trainable}
    print(dt, o)
for n in res["fp32"]:
    a,b = res["fp32"][n], res["bf16"][n]
    cos = (a*b).sum()/np.sqrt((a*a).sum()*(b*b).sum()+1e-300)
    print(f"{n:45s} cos={cos:.4f} |bf16|/|fp32|={np.linalg.norm(b)/ (np.linalg.norm(a)+1e-300):.3f} |fp32|={np.linalg.norm(a):.3e}")
