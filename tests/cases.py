"""Parity cases shared by the golden-fixture generator and the GPU tests."""
from oracle.smct_oracle import Config

CASES = {
    # BASELINE.json configs[0] (README synthetic_model_1): d_model 8, dff 8, 10 particles, bs 32, seq_len 24
    "c1": dict(cfg=Config(B=32, P=10, S=24, D=8, dff=8), sigma2=0.5),
    # configs[1] at reduced batch (P=100)
    "c2_b4": dict(cfg=Config(B=4, P=100, S=24, D=8, dff=8), sigma2=0.5),
    # configs[2] at reduced batch: multivariate F=5, d_model 32, 60 particles
    "c3_b8": dict(cfg=Config(B=8, P=60, S=24, D=32, dff=32, F_x=5, F_y=5), sigma2=0.5),
    # regression-era smoke (pyc __main__): b=8 S=10 F=1 D=6 dff=24 attn_window=4 P=20 sigma^2=0.1 sigma_obs=0.5
    "pyc_smoke": dict(cfg=Config(B=8, P=20, S=10, D=6, dff=24, attn_window=4, noise_z_mode="pyc", sigma_obs=0.5), sigma2=0.1),
    # SMC_TransformerCell.py:208-212 smoke: d_model 12, 3 heads, dff 24, 20 particles
    "heads3": dict(cfg=Config(B=3, P=20, S=9, D=12, H=3, dff=24), sigma2=0.1),
    "window2": dict(cfg=Config(B=2, P=17, S=12, D=16, dff=16, attn_window=2), sigma2=0.3),
    "stop_resampling": dict(cfg=Config(B=3, P=12, S=10, D=8, dff=8, len_resampling=5), sigma2=0.5),
    "attention_only": dict(cfg=Config(B=3, P=9, S=8, D=16, dff=0, full_model=False), sigma2=0.5),
    "multi_sigma": dict(cfg=Config(B=2, P=33, S=7, D=32, dff=16, F_x=3, F_y=2), sigma2=0.2, multi=True),
    # the checkout's own path: Embedding input, classification weights passed as logits
    "classification": dict(cfg=Config(B=4, P=10, S=10, D=16, dff=16, F_x=50, F_y=50, weights="classification",
                                      input_mode="embedding"), sigma2=0.1),
    "no_noise": dict(cfg=Config(B=4, P=1, S=10, D=8, dff=8, noise=False), sigma2=0.5),
    "d64": dict(cfg=Config(B=2, P=70, S=20, D=64, dff=256), sigma2=0.5),
    "d128_h4": dict(cfg=Config(B=2, P=33, S=6, D=128, H=4, dff=64, F_y=3), sigma2=0.5),
    "p1": dict(cfg=Config(B=5, P=1, S=6, D=8, dff=8), sigma2=0.5),
    "s1": dict(cfg=Config(B=2, P=5, S=1, D=8, dff=8), sigma2=0.5),
}

# shapes at the C4 particle count (P=1024, d_model 64, dff 256), short enough for the oracle
BIG_CASES = {
    "c4_b2_s16": dict(cfg=Config(B=2, P=1024, S=16, D=64, dff=256), sigma2=0.5),
}

# shapes served by the fused persistent kernel (d_model 64, dff 64..256, 1 head, full model, gaussian weights)
FUSED_CASES = {
    "f_p70": dict(cfg=Config(B=2, P=70, S=20, D=64, dff=256), sigma2=0.5),
    "f_p1": dict(cfg=Config(B=3, P=1, S=5, D=64, dff=256), sigma2=0.5),
    "f_p128_s9": dict(cfg=Config(B=2, P=128, S=9, D=64, dff=256, F_x=3, F_y=3), sigma2=0.3),
    "f_p129_window": dict(cfg=Config(B=2, P=129, S=12, D=64, dff=256, attn_window=3), sigma2=0.5),
    "f_stop_resampling": dict(cfg=Config(B=2, P=40, S=10, D=64, dff=256, len_resampling=4), sigma2=0.5),
    "f_multi_pyc": dict(cfg=Config(B=2, P=33, S=8, D=64, dff=256, noise_z_mode="pyc", F_y=2), sigma2=0.2, multi=True),
    "f_s1": dict(cfg=Config(B=2, P=16, S=1, D=64, dff=256), sigma2=0.5),
    "f_p300": dict(cfg=Config(B=3, P=300, S=24, D=64, dff=256), sigma2=0.5),
    # dff = 64, 128, 192: one to three FFN chunks of 64 hidden units
    "f_dff64": dict(cfg=Config(B=2, P=50, S=10, D=64, dff=64), sigma2=0.5),
    "f_dff128": dict(cfg=Config(B=2, P=130, S=8, D=64, dff=128, F_x=2, F_y=2), sigma2=0.3),
    "f_dff192": dict(cfg=Config(B=2, P=33, S=9, D=64, dff=192), sigma2=0.5),
    # a long trajectory: many old slots per particle group (shared-unit fast path, several column groups per warp)
    "f_s136": dict(cfg=Config(B=1, P=48, S=136, D=64, dff=128), sigma2=0.5),
}
