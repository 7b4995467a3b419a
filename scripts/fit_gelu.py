"""Fit of the tanh-form GELU used by the head kernels (csrc/mlp_common.cuh): odd polynomial u(x) with tanh(u(x)) ~ erf(x/sqrt 2) on |x| <= 8.
    python scripts/fit_gelu.py   (prints the coefficients and the max errors of the activation and of its derivative)"""
import numpy as np
from scipy.special import erf
from scipy.optimize import minimize
x = np.linspace(-8, 8, 20001)
Phi = 0.5*(1+erf(x/np.sqrt(2)))
phi = np.exp(-x*x/2)/np.sqrt(2*np.pi)
gelu = x*Phi; dgelu = Phi + x*phi
def model(c, x):
    x2 = x*x
    p = c[-1]
    for k in range(len(c)-2, -1, -1): p = p*x2 + c[k]
    u = x*p
    t = np.tanh(u)
    up = 0; # derivative of u: sum (2k+1) c_k x^(2k)
    for k in range(len(c)-1, -1, -1): up = up*x2 + (2*k+1)*c[k]
    a = 0.5*x*(1+t)
    g = 0.5*(1+t) + 0.5*x*(1-t*t)*up
    return a, g
def loss(c):
    a, g = model(c, x)
    return max(np.max(np.abs(a-gelu)), 0.5*np.max(np.abs(g-dgelu)))
for n in (2,3,4):
    c0 = np.array([0.7978845608, 0.0356774081, 0.0, 0.0][:n])
    best = None
    for trial in range(6):
        r = minimize(loss, c0*(1+0.01*np.random.randn(n)) if trial else c0, method='Nelder-Mead', options=dict(xatol=1e-12, fatol=1e-12, maxiter=40000, maxfev=40000))
        if best is None or r.fun < best.fun: best = r
        c0 = best.x
    a, g = model(best.x, x)
    print(n, repr(best.x), 'max|a err|', np.max(np.abs(a-gelu)), 'max|g err|', np.max(np.abs(g-dgelu)))
