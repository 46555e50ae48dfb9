import os, sys, json
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import fr_bytes, rand_fr
from threshold_crypto_b200._lib import Engine
E = Engine(devices=[0]); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(7)
n, t = 1 << 14, 10; m = t + 1
pts = E.sign_g2_batch(rand_fr(rng, n * m), np.tile(E.hash_g2_batch([b"x"])[0], (n * m, 1)))
idx = np.stack([np.sort(rng.choice(32, size=m, replace=False)) for _ in range(n)])
xs = fr_bytes([int(j) + 1 for j in idx.reshape(-1)])
d_x, d_sh = torch.from_numpy(xs).to(dev), torch.from_numpy(pts.reshape(-1)).to(dev)
d_out = torch.zeros(n * 192, dtype=torch.uint8, device=dev); d_st = torch.zeros(n, dtype=torch.uint8, device=dev)
res = {}
for algo in (0, 4):
    E.set_msm_algo(algo)
    for sub in (512, 1024, 2048, 4096, 8192, 16384):
        fn = lambda: E.dev_call("tcb_combine_g2_batch_dev", st, ("size", sub), ("size", t), d_x.data_ptr(), d_sh.data_ptr(), d_out.data_ptr(), d_st.data_ptr())
        fn(); fn(); torch.cuda.synchronize()
        ts = []
        for k in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[f"algo{algo}_n{sub}"] = round(min(ts), 3)
print(json.dumps(res))
