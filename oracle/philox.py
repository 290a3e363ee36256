"""Host restatement of the device counter-based RNG (Philox4x32-7 + Box-Muller).

*** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***  (see oracle/smct_oracle.py header)

The reference draws with ``tf.random.normal`` (self_attention_SMC.py:95) and
``tf.random.categorical`` (SMC_TransformerCell.py:166); neither stream is reproducible outside
TensorFlow, so throughput runs of the CUDA path use their own counter-based generator.  This file
restates that generator so tests can (a) check the device integer stream bit-for-bit (uniforms for
the categorical draw are exact: integer -> fp64), and (b) check the device normals to float tolerance
(device logf/sincosf differ from numpy's in the last ulps; parity tests therefore read the normals
back from the device fill entry point ``smct_fill_noise`` and inject those into the oracle).

Stream layout (must match smc-t-v2_b200/csrc/smct_common.cuh):
  key     = (seed_lo, seed_hi)
  counter = (group_lo, group_hi, stream, t)
  stream  : 0=eps_k 1=eps_q 2=eps_v 3=eps_z 4=uniform
  normals : group = (b_global*P + p) * ceil(D/4) + d//4, element d%4 of the 4 outputs
            (x0,x1)->Box-Muller->(n0,n1), (x2,x3)->(n2,n3); u1 = fl32(x0) 2^-32 + 2^-33, angle from fl32(x1)
  uniform : group = b_global*P + p ; u = ((x0>>5)*2^26 + (x1>>6)) * 2^-53
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


ROUNDS = 7   # = SMCT_PHILOX_ROUNDS (smc-t-v2_b200/csrc/smct_common.cuh): "Philox4x32-7", the smallest Crush-resistant variant


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10: the round function is pinned by the Random123 known-answer vectors (tests/test_oracle_cpu.py)."""
    return philox4x32(c0, c1, c2, c3, k0, k1, rounds=10)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=ROUNDS):
    """Vectorised Philox4x32-R (Salmon et al., SC'11). All inputs uint32 arrays/scalars (broadcast)."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def box_muller(xa, xb):
    """u1 = fl32(xa) 2^-32 + 2^-33 in (0,1] (one fused multiply-add, like the device), angle = fl32(xb) 2pi 2^-32 - pi(1-2^-32);
    r = sqrt(-2 ln u1).  The device evaluates lg2 / sqrt / sin / cos on the special-function unit, so the normals agree to
    float tolerance, not bit for bit."""
    fa = np.asarray(xa, np.uint32).astype(np.float32).astype(np.float64)
    fb = np.asarray(xb, np.uint32).astype(np.float32).astype(np.float64)
    u1 = (fa * float(np.float32(2.0 ** -32)) + float(np.float32(2.0 ** -33))).astype(np.float32)
    th = (fb * float(np.float32(1.4629180792671596e-09)) + float(np.float32(-3.1415926528583237))).astype(np.float32)
    r = np.sqrt(np.float32(-2.0) * np.log(u1, dtype=np.float32), dtype=np.float32)
    return (r * np.cos(th, dtype=np.float32)).astype(np.float32), (r * np.sin(th, dtype=np.float32)).astype(np.float32)


def normals(seed, stream, t, b0, B, P, D):
    """eps[b,p,d] for one stream/timestep; b0 = global index of the first sequence."""
    GD = (D + 3) // 4
    bp = (np.arange(b0, b0 + B, dtype=np.uint64)[:, None] * np.uint64(P) + np.arange(P, dtype=np.uint64)[None, :])
    grp = bp[:, :, None] * np.uint64(GD) + np.arange(GD, dtype=np.uint64)[None, None, :]
    x0, x1, x2, x3 = philox4x32(grp & MASK, grp >> np.uint64(32), np.uint64(stream), np.uint64(t),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    n0, n1 = box_muller(x0, x1)
    n2, n3 = box_muller(x2, x3)
    out = np.stack([n0, n1, n2, n3], axis=-1).reshape(B, P, GD * 4)
    return np.ascontiguousarray(out[:, :, :D])


def uniforms(seed, t, b0, B, P):
    grp = (np.arange(b0, b0 + B, dtype=np.uint64)[:, None] * np.uint64(P) + np.arange(P, dtype=np.uint64)[None, :])
    x0, x1, _, _ = philox4x32(grp & MASK, grp >> np.uint64(32), np.uint64(4), np.uint64(t),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    hi = (x0 >> np.uint32(5)).astype(np.float64)
    lo = (x1 >> np.uint32(6)).astype(np.float64)
    return (hi * 67108864.0 + lo) * (2.0 ** -53)


def draw_all(seed, b0, B, P, S, D):
    eps = np.stack([np.stack([normals(seed, w, t, b0, B, P, D) for t in range(S)]) for w in range(4)])
    u = np.stack([uniforms(seed, t, b0, B, P) for t in range(S)])
    return eps, u
