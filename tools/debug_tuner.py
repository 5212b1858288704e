import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from eeyore_b200.samplers import HMC
from eeyore_b200.tuners import HMCDATuner
from gpu_helpers import dataset, loader, make_model, npy
part = sys.argv[1]
arch, P, C, T, nb, l, e0 = "2321", 20, 21, 40, 25, 0.5, 0.08
rng = np.random.default_rng(4)
theta0 = rng.normal(size=(C, P))
z, u = rng.normal(size=(T, C, P)), rng.uniform(size=(T, C))
m = make_model(arch, "f64", 3 ** 0.5); ds = dataset(arch, "f64")
if part in ("lanes1", "lanes4"):
    s = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0), lanes_per_chain=int(part[-1]))
    s.set_noise_tape(torch.from_numpy(z), torch.from_numpy(u))
    s.run(num_epochs=T, num_burnin_epochs=nb)
    torch.cuda.synchronize(); print(part, "ok", s.step[:4], s.num_steps[:8])
elif part == "philox":
    a = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0), seed=9)
    a.run(num_epochs=T, num_burnin_epochs=nb); torch.cuda.synchronize(); print("philox single ok", a.num_steps[:8])
elif part == "split":
    b = HMC(m, theta0=torch.from_numpy(theta0), dataloader=loader(ds), tuner=HMCDATuner(l=l, e0=e0), seed=9)
    b.run(num_epochs=10, num_burnin_epochs=nb); torch.cuda.synchronize(); print("split 1 ok")
    b.run(num_epochs=T, num_burnin_epochs=nb); torch.cuda.synchronize(); print("split 2 ok")
