"""NumPy restatement of the Monte-Carlo input generator of chirpgp_b200/csrc/cgp_sim.cu (TEST INFRASTRUCTURE ONLY).

The simulation recursion is the reference's (tetralith/jobs/crlb_ekf.py:41-56, test/test_crlb.py:41-55, tools.py:81-170):
    x_0 = m0 + chol(P0) eps,   x_k = mean(x_{k-1}) + chol(Sigma) eps_k,   y_k = H x_k + sqrt(Xi) eps'_k.
The reference draws eps with jax.random.normal (threefry), which cannot be reproduced without JAX; the kernel uses the
counter-based Philox4x32-10 generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
Random123 known-answer vectors below) + Box-Muller.  This file restates exactly that generator so that the kernel can be
checked sample by sample."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# Random123 kat_vectors: philox4x32-10  counter / key -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values); key words are Python ints."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    ka, kb = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(ka), lo1, hi0 ^ c[3] ^ np.uint64(kb), lo0]
        ka, kb = (ka + W0) & 0xFFFFFFFF, (kb + W1) & 0xFFFFFFFF
    return c


def normal2(seed, draw, step, traj):
    """Two standard normals per (draw, step, trajectory): 53-bit uniforms u1 in (0, 1], u2 in [0, 1), Box-Muller."""
    traj = np.asarray(traj, dtype=np.uint64)
    c = philox4x32_10(draw, step, traj & MASK, traj >> np.uint64(32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = ((c[0] << np.uint64(32)) | c[1]) >> np.uint64(11)
    b = ((c[2] << np.uint64(32)) | c[3]) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 1.) * 2. ** -53
    u2 = b.astype(np.float64) * 2. ** -53
    r = np.sqrt(-2. * np.log(u1))
    return r * np.cos(2. * np.pi * u2), r * np.sin(2. * np.pi * u2)


def normals(seed, n, step, traj):
    """eps[..., 0..n): draws 0, 1, ... of (step, trajectory)."""
    out = []
    for i in range(0, n, 2):
        z0, z1 = normal2(seed, i // 2, step, traj)
        out += [z0, z1]
    return np.stack(out[:n], axis=-1)


def chol_psd(A):
    d = A.shape[0]
    L = np.zeros((d, d))
    for j in range(d):
        s = A[j, j] - L[j, :j] @ L[j, :j]
        L[j, j] = np.sqrt(s) if s > 0. else 0.
        for i in range(j + 1, d):
            t = A[i, j] - L[i, :j] @ L[j, :j]
            L[i, j] = t / L[j, j] if L[j, j] > 0. else 0.
    return L


def simulate(mean_fn, Sigma, H, Xi, m0, P0, T, B, seed, first_trajectory=0):
    """mean_fn(x (B, d)) -> (B, d).  Returns x0 (B, d), xs (B, T, d), ys (B, T)."""
    d = m0.shape[-1]
    traj = np.arange(B, dtype=np.uint64) + np.uint64(first_trajectory)
    L0, Ls = chol_psd(np.asarray(P0)), chol_psd(np.asarray(Sigma))
    x = m0 + normals(seed, d, 0, traj) @ L0.T
    x0 = x.copy()
    xs, ys = np.empty((B, T, d)), np.empty((B, T))
    for t in range(T):
        eps = normals(seed, d + 1, t + 1, traj)
        x = mean_fn(x) + eps[:, :d] @ Ls.T
        xs[:, t] = x
        ys[:, t] = x @ H + np.sqrt(Xi) * eps[:, d]
    return x0, xs, ys
