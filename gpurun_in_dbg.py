import numpy as np, sys
sys.path.insert(0,'.')
from crime_b200 import GetHI, params_from_tables
from oracle.binding import Oracle
o=Oracle()
tabs=dict(np.load('tests/golden/ref_tables_nu150.npz'))
for nside in (16,1024):
    p = params_from_tables(tabs, n_grid=32, n_side=nside)
    rng = np.random.default_rng(nside)
    n = 2_000_000
    r = rng.uniform(0.2 * float(tabs["r_min"]), 1.3 * float(tabs["r_max"]), n)
    u = rng.standard_normal((n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    pos = u * r[:, None]
    k = n // 10
    pos[:k, 2] = np.sign(pos[:k, 2]) * np.abs(pos[:k, 0]) * rng.uniform(50, 5000, k)
    pos[k:2 * k, 2] = np.hypot(pos[k:2 * k, 0], pos[k:2 * k, 1]) * (2 / 3) / np.sqrt(1 - 4 / 9) * rng.choice([-1, 1], k) * (1 + rng.uniform(-1e-12, 1e-12, k))
    pos[2 * k:3 * k, 1] = rng.uniform(-1e-9, 1e-9, k)
    dz = rng.normal(0, 2e-3, n)
    with GetHI(p) as g:
        sh, px = g.points_to_shell_pixel(pos, dz)
    sh_o, px_o = o.points_to_shell_pixel(p, pos, dz)
    bad = np.nonzero(px != px_o)[0]
    print('nside',nside,'mismatches',len(bad),'shell mism',(sh!=sh_o).sum())
    for i in bad[:12]:
        x,y,z=pos[i]; rr=np.sqrt(x*x+y*y+z*z)
        print(i, 'group', i//k, repr(x),repr(y),repr(z),'cth',z/rr,'phi',np.arctan2(y,x),'gpu',px[i],'cpu',px_o[i],'sh',sh[i])
