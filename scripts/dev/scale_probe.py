import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import smct_oracle as O
from tests import util
cfg = O.Config(B=2, P=48, S=12, D=64, dff=256, F_x=2, F_y=2)
for scale in (2.0, 4.0, 8.0):
    p, x_in, y, eps, u = util.make_case(cfg, seed=17, sigma2=0.5)
    for k in ("Wq", "Wk", "Wv", "bq", "bk", "bv", "Wz", "W1", "W2"):
        p[k] = (np.asarray(p[k], np.float32) * np.float32(scale)).astype(np.float32)
    sig = float(scale) ** 2 * 0.5
    for k in ("logvar_k", "logvar_q", "logvar_v", "logvar_z"):
        p[k] = np.full_like(np.asarray(p[k], np.float32), np.log(sig), dtype=np.float32)
    u, ora = util.screen_uniforms(cfg, p, x_in, y, eps, u)
    for algo in ("staged", "fused_ffma", "fused"):
        dev = util.run_device(cfg, p, x_in, y, eps, u, i_forced=ora["i_t"], want=("i_out","logw","loss_sums","noise_z","noise_q"), algo=algo)
        errs={}
        for k in ("pred","K","R","w","noise_z"):
            a=dev[k].astype(np.float64); b=ora[k].astype(np.float64)
            errs[k]=float(np.abs(a-b).max()/np.abs(b).max())
            tol=1e-4*np.abs(b)+1e-5*np.abs(b).max()
            errs[k+"_bad"]=int((np.abs(a-b)>tol).sum())
        print(scale, algo, {k:"%.1e"%v for k,v in errs.items()})
