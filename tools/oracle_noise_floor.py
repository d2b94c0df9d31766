"""Round-off noise floor of the reference ALGORITHM at BASELINE cfg 2 full length (TRAPPIST-1, h = 0.06 d, 1600 d = 26,667 steps):
the same oracle source compiled without and with FMA contraction.  Output kept in profiles/r01_oracle_noise_floor.txt; the full-length
GPU parity test uses it as the tolerance for Jacobian-type outputs (transit times and x, v keep 1e-11)."""
import sys, numpy as np, time
sys.path[:0]=['/root/repo','/root/repo/nbodygradient.jl_b200']
from oracle.binding import Oracle, build
build(); 
el=np.loadtxt('/root/repo/tests/golden/elements.txt',delimiter=',')
def run(o):
    x,v,jac=o.init_nbody(el,7257.0)
    s=o.new_state(x,v,el[:,0],7257.0)
    t=time.time(); r=o.transit_timing(s,0.06,1600.0,1062,grad=True,jac_init=jac); print('time',time.time()-t)
    return s,r
a_s,a=run(Oracle()); b_s,b=run(Oracle(fast=True))
def rel(p,q): return np.max(np.abs(p-q))/np.max(np.abs(q))
print('counts equal',np.array_equal(a['count'],b['count']), a['count'].sum())
m=a['tt']!=0
print('tt rel',np.max(np.abs(a['tt'][m]-b['tt'][m])/np.abs(a['tt'][m])))
print('dtdq0 rel',rel(a['dtdq0'],b['dtdq0']),'dtdelements rel',rel(a['dtdelements'],b['dtdelements']))
print('x rel',rel(a_s['x'],b_s['x']),'v rel',rel(a_s['v'],b_s['v']),'jac rel',rel(a_s['jac_step_cm'],b_s['jac_step_cm']))
