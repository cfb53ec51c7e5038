"""Interpolation error of the lattice resampler's far field as a function of lattice spacing, cut-off radius and the
smoothness of the blending polynomial (numpy fp64, one control point of unit weight on the bench canvas 755 x 1783).

Why: round 2 tried a TWO-LEVEL evaluation of the lattice nodes (all 63 exact terms on a lattice of spacing 2h with cut-off
2R, the fine lattice by quintic interpolation + difference terms; 74 -> 48 us for the node kernels).  The measured coordinate
error went from mean 1.9e-4 / max 1.6e-3 px to mean 4.9e-4 / max 5.4e-3 px (the reference's own fp32 evaluation: 2.0e-4 /
2.1e-3), so the test that ours must not be worse than the reference failed.  This script shows why: the far field is only
C^3 across the circle s = R^2 (cubic Taylor blend), so the quintic interpolation error is the O(h^4) jump term, which scales
like R^2 (h / R)^4: a coarse level with (2h, kR) has 16 / k^2 times the error of the fine level (4x at k = 2, as measured;
k >= 8 would be needed, i.e. no near list at all), and higher-order blends trade the jump for a badly resolved polynomial.
The two-level code was removed; the one-level error (1.8e-3 px per unit weight at the maximum) is what bounds the accuracy of
the lattice path.

    python profiles/lattice_error_sim.py
"""
import numpy as np
from numpy.polynomial import polynomial as Pn
Wo,Ho=1783,755
ux,uy=2.0/(Wo-1),2.0/(Ho-1)
SX,SY=16,6
h=max(SX*ux,SY*uy); R=4.3*h
eps=1e-6
def phi(s): return s*np.log(s+eps)
def taylor(a,deg):
    # derivatives of phi at a
    L=np.log(a+eps); d=[a*L, L+a/(a+eps), 1/(a+eps)+eps/(a+eps)**2, -1/(a+eps)**2-2*eps/(a+eps)**3,
       2/(a+eps)**3+6*eps/(a+eps)**4, -6/(a+eps)**4-24*eps/(a+eps)**5, 24/(a+eps)**5+120*eps/(a+eps)**6, -120/(a+eps)**6]
    import math
    return [d[k]/math.factorial(k) for k in range(deg+1)]
def far(s,a,deg):
    c=taylor(a,deg); u=s-a
    p=np.zeros_like(s)
    for k in reversed(range(deg+1)): p=p*u+c[k]
    return np.where(s>=a, phi(s), p)
def lag_w(t):
    xs=np.array([-2,-1,0,1,2,3.]); w=np.ones(6)
    for j in range(6):
        for m in range(6):
            if m!=j: w[j]*=(t-xs[m])/(xs[j]-xs[m])
    return w
def interp_err(sx,sy,a,deg,cx,cy,region=None):
    # lattice with spacing (sx,sy) pixels; evaluate far field at nodes; interpolate to every pixel in a window around the control point; return max abs err (normalised units)
    # window: +-3R around control point
    Rpx_x=int(np.sqrt(a)/ux*1.6)+2*sx; Rpx_y=int(np.sqrt(a)/uy*1.6)+2*sy
    px0=int((cx+1)/ux); py0=int((cy+1)/uy)
    xs=np.arange(px0-Rpx_x,px0+Rpx_x); ys=np.arange(py0-Rpx_y,py0+Rpx_y)
    # nodes
    nx0=(xs[0]//sx)-2; nx1=(xs[-1]//sx)+4; ny0=(ys[0]//sy)-2; ny1=(ys[-1]//sy)+4
    gx=np.arange(nx0,nx1+1)*sx*ux-1; gy=np.arange(ny0,ny1+1)*sy*uy-1
    S=(gx[None,:]-cx)**2+(gy[:,None]-cy)**2
    F=far(S,a,deg)
    # interpolate
    out=np.zeros((len(ys),len(xs)))
    wxs=[lag_w(k/sx) for k in range(sx)]; wys=[lag_w(k/sy) for k in range(sy)]
    # y contraction then x
    for iy,y in enumerate(ys):
        cyi=y//sy; wy=wys[y-cyi*sy]
        rowv=(wy[:,None]*F[cyi-2-ny0:cyi+4-ny0,:]).sum(0)
        for ix,x in enumerate(xs):
            cxi=x//sx; wx=wxs[x-cxi*sx]
            out[iy,ix]=(wx*rowv[cxi-2-nx0:cxi+4-nx0]).sum()
    X=xs*ux-1; Y=ys*uy-1
    Sx=(X[None,:]-cx)**2+(Y[:,None]-cy)**2
    ex=far(Sx,a,deg)
    e=np.abs(out-ex)
    return e.max(), e.mean()
rng=np.random.default_rng(0)
for (sx,sy,a,deg,name) in [(SX,SY,R*R,3,'fine cubic R'),(2*SX,2*SY,4*R*R,3,'coarse cubic 2R'),(2*SX,2*SY,4*R*R,5,'coarse quintic 2R'),(2*SX,2*SY,4*R*R,4,'coarse quartic 2R'),(2*SX,2*SY,4*R*R,7,'coarse deg7 2R'),(2*SX,2*SY,9*R*R,3,'coarse cubic 3R'),(2*SX,2*SY,9*R*R,5,'coarse quintic 3R')]:
    mx=[];mn=[]
    for t in range(3):
        cx,cy=rng.uniform(-0.3,0.3,2)
        a_,b_=interp_err(sx,sy,a,deg,cx,cy); mx.append(a_); mn.append(b_)
    print('%-20s max err*640 = %.2e px per unit weight, mean (window) %.2e'%(name,max(mx)*640,np.mean(mn)*640))
