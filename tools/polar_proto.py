"""CPU prototype: third-order remainder of the polar-cap expansion of V = nside sqrt(3 (1 - |z|/r)) used by gh_group_math.cuh;
prints the largest observed error as a multiple of the candidate bounds (the kernel uses 0.45 ns (d/rho)^3 (rho/r) + 0.4 ns (d/r)^3).
Not part of the product."""
import numpy as np
rng=np.random.default_rng(3)
def F_exact(x,y,z,ns):
    r=np.sqrt(x*x+y*y+z*z); rho2=x*x+y*y
    return ns*np.sqrt(3.0)*np.sqrt(rho2/(r*(r+np.abs(z))))
def coeffs_F(x,y,z,ns):
    # u = 1 - s z/r ; F = K sqrt(u)
    s=np.sign(z); r2=x*x+y*y+z*z; r=np.sqrt(r2); ir=1/r; ir3=ir**3; ir5=ir**5; rho2=x*x+y*y
    u=rho2/(r*(r+np.abs(z)))
    # g = grad(z/r), H = hess(z/r)
    gx,gy,gz=-x*z*ir3,-y*z*ir3,rho2*ir3
    Hxx=-z*ir3+3*z*x*x*ir5; Hyy=-z*ir3+3*z*y*y*ir5; Hzz=-3*z*ir3+3*z**3*ir5
    Hxy=3*z*x*y*ir5; Hxz=-x*ir3+3*x*z*z*ir5; Hyz=-y*ir3+3*y*z*z*ir5
    K=ns*np.sqrt(3.0); su=np.sqrt(u)
    a=-s*K/(2*su); b=-K/(4*u*su)
    c=dict(c=K*su,x=a*gx,y=a*gy,z=a*gz,
           xx=0.5*(a*Hxx+b*gx*gx),yy=0.5*(a*Hyy+b*gy*gy),zz=0.5*(a*Hzz+b*gz*gz),
           xy=a*Hxy+b*gx*gy,xz=a*Hxz+b*gx*gz,yz=a*Hyz+b*gy*gz)
    return c
n=400000; ns=256
for dcell in (17.45, 4.4):
  for rho_cells_min in (15,30,60):
    r=rng.uniform(1300,4500,n); cth=rng.uniform(0.667,0.9999,n)*rng.choice([-1,1],n)
    ph=rng.uniform(0.05,np.pi/2-0.05,n)+rng.integers(0,4,n)*np.pi/2
    st=np.sqrt(1-cth*cth); x,y,z=r*st*np.cos(ph),r*st*np.sin(ph),r*cth
    rho=np.sqrt(x*x+y*y)
    m=rho>rho_cells_min*dcell
    x,y,z,r,rho=x[m],y[m],z[m],r[m],rho[m]
    o=2*dcell*(rng.random((3,len(x)))-0.5)
    # extreme corners too
    o=np.where(rng.random(o.shape)<0.5,np.sign(o)*dcell,o)
    c=coeffs_F(x,y,z,ns)
    Ft=(c['c']+c['x']*o[0]+c['y']*o[1]+c['z']*o[2]+c['xx']*o[0]**2+c['yy']*o[1]**2+c['zz']*o[2]**2+c['xy']*o[0]*o[1]+c['xz']*o[0]*o[2]+c['yz']*o[1]*o[2])
    Fe=F_exact(x+o[0],y+o[1],z+o[2],ns)
    d=np.sqrt((o**2).sum(0))
    err=np.abs(Ft-Fe)
    b1=ns*(d/rho)**3*(rho/r)
    b2=ns*(d/r)**3
    print(dcell,rho_cells_min,'max err px',err.max(),'ratio to ns(d/rho)^3(rho/r):',(err/b1).max(),' ratio incl (d/r)^3 term:',(err/(b1+b2)).max(), 'ratio (d/rho)^3 only', (err/(ns*(d/rho)**3)).max())
