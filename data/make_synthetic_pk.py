"""Generate the synthetic linear P(k) used by tests and bench (no network, and the reference's CAMB
table is not copied into this repo).  BBKS (Bardeen et al. 1986) transfer function with Sugiyama's
shape parameter; CAMB-style two-column text: k [h/Mpc], P(k) [(Mpc/h)^3], log-spaced like the
reference's data/Pk_CAMB_test.dat (807 rows over 1e-4..1e3 h/Mpc, 5 significant digits).
The amplitude is arbitrary: GetHI renormalises to sigma_8 (reference src/cosmo.c:331-338).
"""
import numpy as np


def bbks_pk(k, om=0.3, ob=0.049, h=0.67, ns=0.96):
    gamma = om * h * np.exp(-ob * (1 + np.sqrt(2 * h) / om))
    q = k / gamma
    t = np.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    return k ** ns * t ** 2


def main(path="Pk_synth.dat", n=807):
    k = np.logspace(-4, 3, n)
    pk = bbks_pk(k)
    pk *= 2.0e4 / pk.max()
    with open(path, "w") as f:
        for a, b in zip(k, pk):
            f.write(f"{a:15.5E}{b:15.5E}\n")


if __name__ == "__main__":
    import sys
    main(*sys.argv[1:2])
